// kernels.h -- launchers of the sm_100a kernels (kernels.cu).  All pointers are DEVICE
// pointers; every launcher enqueues on `s` and returns the launch status.
#ifndef MALLIE_B200_KERNELS_H_
#define MALLIE_B200_KERNELS_H_

#include <cuda_runtime_api.h>

#include <vector>

#include "layout.h"
#include "mallie_b200.h"

namespace mb200 {

// Optional per-kernel timing: CUDA events recorded on the launching stream around every kernel of a
// class (bench.py's roofline leg: the duration of the traversal kernels inside the timed region).
enum KernelClass { kKCameraTrace = 0, kKShadowTrace, kKBounceTrace, kKShade, kKResolve, kKQueryTrace, kKClasses };
struct KernelTimer {
  bool enabled = false;
  struct Span {
    int cls;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;        // recorded, not yet collected
  std::vector<cudaEvent_t> pool;  // recycled events
  void begin(int cls, cudaStream_t s);
  void end(cudaStream_t s);
  // After the streams have been synchronised: adds, per class, the length of the UNION of its spans (kernels of
  // consecutive batches run on two streams and may overlap, see FramePipe) and the launch counts; *trace_union_ms
  // (nullable) receives the union over the three traversal classes.  Recycles the events.
  void collect(double ms[kKClasses], unsigned long long launches[kKClasses], double *trace_union_ms = nullptr);
  void release();
};

// Second stream + events of the frame pipeline: consecutive batches of a frame alternate between the scene's
// stream and `aux`, so that the drain phase of one persistent traversal kernel (the last, longest rays of a
// launch run at single-warp latency while the rest of the GPU idles: ~0.35 ms per launch on the 1 M-triangle
// scene) is filled by the next batch's kernel instead of being waited for.
//
// It also owns the longest-rays-first schedule.  What is left of the tail belongs to the last launch of a
// frame, whose long rays (silhouette tiles) happen to be handed out late.  The traversal kernels flag the tiles
// that held a ray of more than kHotSteps steps (`hot`); the next frame over the same tile layout hands those
// tiles out first (`order`, a stable partition of the tile indices built by k_build_order), so its long rays
// start while the GPU is full.  Pure scheduling: which warp traces a ray never changes a result.
// When do which rows of a frame become final?  launch_frame cuts a frame into batches by tile rows (every batch holds all
// passes of its rows), so a caller that copies the frame somewhere -- to the host, to the other ranks -- can start with the
// first rows while the later ones are still traced.  Rows are those of the caller's buffer (local rows of a compact band
// buffer, image rows otherwise).  A chunk is final once BOTH of its events have completed (the frame's batches alternate
// between two streams; the second event is null when one stream carried the whole chunk).  n == 0: the frame was not
// cut by rows (passes did not fit a batch, or non-compact bands): everything is final when the stream is.
constexpr int kMaxFrameChunks = 8;
struct FrameChunks {
  int want = 2;  // in: chunks the caller would like (>= 1)
  int n = 0;     // out
  int row0[kMaxFrameChunks], row1[kMaxFrameChunks];
  cudaEvent_t done_a[kMaxFrameChunks], done_b[kMaxFrameChunks]; // owned by the FramePipe
};

struct FramePipe {
  cudaStream_t aux = nullptr;
  cudaStream_t copy = nullptr; // device -> host / collective work that overlaps the frame's later batches
  cudaEvent_t fork = nullptr, join = nullptr, resolved[2] = {nullptr, nullptr};
  cudaEvent_t chunk_ev[2 * kMaxFrameChunks] = {nullptr};
  unsigned char *hot = nullptr; // [tiles_cap]
  uint32_t *order = nullptr;    // [tiles_cap]
  size_t tiles_cap = 0;
  long long layout[12] = {0};   // the tile layout `hot` was collected for
  bool have_hot = false;
  cudaError_t init();
  cudaError_t reserve_tiles(size_t tiles, cudaStream_t s);
  void release();
};

// `work` is an 8-byte device scratch word (persistent-warp work counter, reset by the launcher);
// `counters` (nullable) is unsigned long long[4]: nodes, tris, rays, max stack (accumulated).
cudaError_t launch_trace_closest(const SceneView &sc, int stack_cap, const mb200_ray *rays, size_t n,
                                 mb200_hit *hits, unsigned long long *work, unsigned long long *counters,
                                 cudaStream_t s, KernelTimer *timer = nullptr);
cudaError_t launch_trace_occluded(const SceneView &sc, int stack_cap, const mb200_ray *rays, const double *tmax,
                                  size_t n, unsigned char *occluded, unsigned long long *work,
                                  unsigned long long *counters, cudaStream_t s, KernelTimer *timer = nullptr);
// K3: BuildIntersection for every ray from its 32-byte hit record (mask nullable).
cudaError_t launch_build_isects(const SceneView &sc, const mb200_ray *rays, const mb200_hit *hits, size_t n,
                                mb200_isect *isects, unsigned char *mask, cudaStream_t s);
cudaError_t launch_generate_rays(const mb200_camera_frame &f, const double *px, const double *py, size_t n,
                                 mb200_ray *rays, cudaStream_t s);
cudaError_t launch_generate_rays_env(const double origin[3], int width, int height, int stereo, const double *px,
                                     const double *py, size_t n, mb200_ray *rays, cudaStream_t s);
cudaError_t launch_generate_grid(const mb200_camera_frame &f, int x0, int y0, int w, int h, mb200_ray *rays,
                                 cudaStream_t s);

// Device scratch of the frame pipeline (hit records, ray queues, per-sample contributions, path state,
// counters); owned by the scene, grown on demand by launch_frame.
struct FrameScratch {
  void *base = nullptr;
  size_t bytes = 0;
};
cudaError_t frame_scratch_reserve(FrameScratch &fs, size_t bytes, cudaStream_t s);
void frame_scratch_release(FrameScratch &fs);

// One frame of num_passes samples per pixel (wavefront: primary trace -> shade -> secondary trace ... ->
// resolve), batched so that a batch's buffers stay within a fixed budget.
// mode 0: image = last pass, count += passes (one pass: render.cc:673-679)
// mode 1: image += passes, count += passes    (AccumImage, main_sdl.cc:138-143)
// mode 2: image = sum of passes, count = passes (fresh frame; nothing read)
// stats: device unsigned long long[8] primary, bounce, shadow, zombie, then (fused primary+shadow / primary-only
// frames) camera nodes, camera triangles, shadow nodes, shadow triangles tested (accumulated).
cudaError_t launch_frame(const SceneView &sc, int stack_cap, const mb200_render_params &p, int num_passes, int mode,
                         float *image, int *count, FrameScratch &scratch, unsigned long long *stats, cudaStream_t s,
                         KernelTimer *timer = nullptr, FramePipe *pipe = nullptr, FrameChunks *chunks = nullptr);

// Accumulated frame + counts -> 8-bit pixels (mode 0: HDRToLDR RGB8, mode 1: Display BGRA8 with gamma 2.2).
cudaError_t launch_resolve_ldr(const float *image, const int *count, size_t npix, int mode, unsigned char *out, cudaStream_t s);

// The eight octant copies of the pair nodes (layout.h: nodes_oct); dst holds 8 * n nodes.
cudaError_t launch_octant_nodes(const PairNode *src, size_t n, PairNode *dst, cudaStream_t s);
// 64-byte copies of the pair nodes (development variant; *bad != 0 afterwards: not representable, do not use them).
cudaError_t launch_pack_nodes64(const PairNode *src, size_t n, PairNode64 *dst, int *bad, cudaStream_t s);
// Traversal copy of the triangle records in a padded layout (layout.h: TriKind kTriF32x64 / kTriF64x96).
cudaError_t launch_pad_tris(const void *src, int src_f32, size_t n, int kind, void *dst, cudaStream_t s);

// Rows owned by one band index (see mb200_render_params::band_rows).
int band_rows_owned(int rows, int band_rows, int count, int index);
// Number of kernel launches issued by this library in this process (bench.py's gpu_launches).
int launches_issued();
void note_launch(); // a kernel of this library launched outside kernels.cu (gather.cu)

} // namespace mb200

#endif

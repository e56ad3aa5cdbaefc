// scene.h -- the device-resident scene behind the opaque mb200_scene handle.
#ifndef MALLIE_B200_SCENE_H_
#define MALLIE_B200_SCENE_H_

#include <cuda_runtime_api.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "kernels.h"
#include "layout.h"
#include "mallie_b200.h"

struct mb200_scene {
  int device = 0;
  cudaStream_t stream = nullptr;
  mb200::SceneView view{};
  int stack_cap = 64;          // traversal stack entries a ray can need (tree depth + 2)
  int tree_depth = 0;
  size_t device_bytes = 0;
  double root_bmin[3] = {0, 0, 0}, root_bmax[3] = {0, 0, 0};

  // owned device allocations
  std::vector<void *> allocs;
  unsigned long long *d_work = nullptr;     // persistent-warp work counter
  unsigned long long *d_counters = nullptr; // [4] traversal counters / render stats

  // staging for host-pointer calls: pinned host + device mirrors, grown on demand
  struct Staging {
    void *pinned = nullptr;
    void *dev = nullptr;
    size_t cap = 0;
  };
  Staging in0, in1, out0, out1;

  // device scratch: the frame wavefront's buffers, and hit records of mb200_trace_closest_full
  mb200::FrameScratch frame_scratch, hit_scratch;

  // per-kernel timing (mb200_scene_timing / mb200_scene_kernel_times)
  mb200::KernelTimer timer;
  // second stream of the frame pipeline (kernels.h)
  mb200::FramePipe pipe;
};

namespace mb200 {

// Validates the reference-layout BVH, re-lays it out (layout.h) and uploads everything.
// On failure returns a negative mb200_status and fills err.
int scene_create(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                 size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                 const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices,
                 std::string *err);
void scene_destroy(mb200_scene *s);
int scene_clone(mb200_scene **out, mb200_scene *src, int device, std::string *err);
// Pieces shared with the device-side build (bvh_build_gpu.cu): device checks + handle + stream; work counters + sync;
// which triangle record the vertices allow.
int scene_open(mb200_scene **out, int device, std::string *err);
int scene_finish(mb200_scene *s, std::string *err);
bool choose_tri_f32(const double *vertices, size_t count);

// Uninitialised host array (std::vector would zero-fill tens of MB on one thread before the parallel loops that
// write every byte touch them).
template <class T> struct RawArray {
  T *p = nullptr;
  size_t n = 0;
  RawArray() = default;
  RawArray(const RawArray &) = delete;
  RawArray &operator=(const RawArray &) = delete;
  RawArray(RawArray &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
  RawArray &operator=(RawArray &&o) noexcept {
    if (this != &o) {
      free(p);
      p = o.p, n = o.n;
      o.p = nullptr, o.n = 0;
    }
    return *this;
  }
  ~RawArray() { free(p); }
  bool resize(size_t k) {
    free(p);
    p = k ? static_cast<T *>(aligned_alloc(128, ((k * sizeof(T) + 127) / 128) * 128)) : nullptr;
    n = p ? k : 0;
    return k == 0 || p != nullptr;
  }
  T *data() { return p; }
  const T *data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

// Host-side relayout only (no CUDA): exposed for CPU tests of the layout logic.
struct Relayout {
  RawArray<PairNode> pairs;
  RawArray<TriRecordF32> tris32;
  RawArray<TriRecordF64> tris64;
  bool f32 = false;
  uint32_t root_ref = 0, root_cnt = 0;
  int depth = 0;
  bool empty = true;
};
int relayout_bvh(Relayout &out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                 const uint32_t *material_ids, const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices,
                 size_t nindices, std::string *err);

} // namespace mb200

#endif

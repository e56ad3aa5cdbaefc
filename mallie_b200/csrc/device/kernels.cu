// kernels.cu -- sm_100a kernels of the Mallie render hot path and their launchers.
//
//   K1 raygen        Camera::GenerateRay           camera.cc:222-240     (k_generate_*; fused into K2 for frames)
//   K2 closest hit   BVHAccel::Traverse            bvh_accel.cc:773-844  (k_trace_sm, trace_sm.cuh)
//   K3 hit record    BuildIntersection             bvh_accel.cc:699-769  (k_build_isects; inside the shade kernels)
//   K4 occlusion     closest-hit t < tmax          (render.cc:425-426 leaves NEE empty)
//   K5 frame         Render/PathTrace + Plane      render.cc:381-456,593-708, prim-plane.cc:8-44
//                    as a wavefront: trace camera rays -> shade -> trace queued rays -> ... -> resolve
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
// (no FMA contraction: see traverse.cuh).
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "shade.cuh"
#include "trace_sm.cuh"
#ifdef MB200_DEV_VARIANTS // measured-and-rejected machines, development builds only
#include "trace_ds.cuh"
#include "trace_ps.cuh"
#include "trace_tr.cuh"
#endif
#include "traverse.cuh"

namespace mb200 {

namespace {

#ifndef MB200_BLOCK
#define MB200_BLOCK 128
#endif
constexpr int kBlock = MB200_BLOCK; // threads per CTA of the traversal kernels

// production configuration of the traversal state machine (A/B history: DESIGN.md §5)
constexpr int kRefillMin = 8;   // idle lanes that trigger a refill
constexpr int kShadeMin = 8;    // parked lanes that trigger a shade step (fused frames)
constexpr int kSmemStack = 12;  // stack entries per thread kept in shared memory (24 KB per CTA)
constexpr int kSmemStackQuery = 8;  // plain closest-hit / any-hit queries over a caller's ray buffer
constexpr int kSmemStackFused = 10; // fused frames: two more 16-byte units per thread hold the lane slot
constexpr int kMinBlocks = 1024 / kBlock;   // resident CTAs per SM the register allocation targets (64 registers)
constexpr unsigned kChunk = 32; // ray indices per atomicAdd
constexpr int kVar = kVarOctNodes; // octant copies of the pair nodes: no sign selects in the inner step (trace_sm.cuh)

__device__ __forceinline__ unsigned int lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------
// K1: ray generation into a ray buffer
// ---------------------------------------------------------------------------
__device__ __forceinline__ void store_ray(mb200_ray *dst, const mb200_camera_frame &f, double dx, double dy, double dz) {
  double2 *o = reinterpret_cast<double2 *>(dst);
  o[0] = make_double2(f.origin[0], f.origin[1]);
  o[1] = make_double2(f.origin[2], dx);
  o[2] = make_double2(dy, dz);
}

__global__ void __launch_bounds__(256) k_generate_rays(const __grid_constant__ mb200_camera_frame frame,
                                                       const double *__restrict__ px, const double *__restrict__ py,
                                                       size_t n, mb200_ray *__restrict__ rays) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double dx, dy, dz;
    generate_ray(frame, px[i], py[i], dx, dy, dz);
    store_ray(rays + i, frame, dx, dy, dz);
  }
}

__global__ void __launch_bounds__(256) k_generate_grid(const __grid_constant__ mb200_camera_frame frame, int x0,
                                                       int y0, int w, int h, mb200_ray *__restrict__ rays) {
  const size_t n = (size_t)w * h;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = x0 + (int)(i % w), y = y0 + (int)(i / w);
    double dx, dy, dz;
    generate_ray(frame, (double)x, (double)y, dx, dy, dz);
    store_ray(rays + i, frame, dx, dy, dz);
  }
}

__global__ void __launch_bounds__(256) k_generate_rays_env(const __grid_constant__ mb200_camera_frame frame, int width,
                                                           int height, int stereo, const double *__restrict__ px,
                                                           const double *__restrict__ py, size_t n,
                                                           mb200_ray *__restrict__ rays) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double ox, oy, oz, dx, dy, dz;
    generate_env_ray(frame.origin, width, height, px[i], py[i], stereo != 0, ox, oy, oz, dx, dy, dz);
    double2 *o = reinterpret_cast<double2 *>(rays + i);
    o[0] = make_double2(ox, oy);
    o[1] = make_double2(oz, dx);
    o[2] = make_double2(dy, dz);
  }
}

// ---------------------------------------------------------------------------
// K2 / K4: the persistent-warp traversal state machine (trace_sm.cuh) as a kernel.
// n_dev (nullable): the ray count lives in device memory (a queue filled by the previous kernel).
// ---------------------------------------------------------------------------
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int SHADE_MIN, int S, int MINB, unsigned CHUNK, int VAR>
__global__ void __launch_bounds__(kBlock, MINB)
    k_trace_sm(const __grid_constant__ SceneView sc, const __grid_constant__ IO io, unsigned long long n,
               const unsigned int *__restrict__ n_dev, unsigned long long *__restrict__ work,
               unsigned long long *__restrict__ gcounters) {
  extern __shared__ uint4 smem_stack[]; // [S stack entries (+ 2 lane-slot units)][kBlock], column layout
  TravStack<S, CAP> st;
  st.init(smem_stack + threadIdx.x, kBlock);
  LaneSlot slot;
  slot.addr0 = st.sm_addr + (uint32_t)S * st.stride_bytes;
  slot.addr1 = slot.addr0 + st.stride_bytes;
  uint32_t top_table = 0;
  if (VAR & kVarTopSmem) { // [stack][lane slots][16 B barrier][top-of-tree table]
    const uint32_t mbar = st.sm_addr - threadIdx.x * 16u + (uint32_t)(S + (IO::kFused ? 2 : 0)) * st.stride_bytes;
    top_table = mbar + 16u;
    if (sc.top_count) stage_top_nodes(sc, mbar, top_table);
  }
  if (n_dev) n = __ldg(n_dev);
  trace_state_machine<IO, TRI, S, CAP, ANYHIT, COUNT, REFILL_MIN, SHADE_MIN, CHUNK, VAR>(sc, io, n, work, st, slot, gcounters,
                                                                                       top_table);
}

#ifdef MB200_DEV_VARIANTS
// The two-rays-per-lane machine (trace_tr.cuh): the rays' read-only halves in dynamic shared memory.
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int HYST, int MINB, unsigned CHUNK>
__global__ void __launch_bounds__(kBlock, MINB)
    k_trace_tr(const __grid_constant__ SceneView sc, const __grid_constant__ IO io, unsigned long long n,
               const unsigned int *__restrict__ n_dev, unsigned long long *__restrict__ work,
               unsigned long long *__restrict__ gcounters) {
  extern __shared__ uint4 smem_fat[]; // [2 rays x kFatUnits][kBlock], column layout
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_fat + threadIdx.x);
  if (n_dev) n = __ldg(n_dev);
  trace_two_ray_machine<IO, TRI, CAP, ANYHIT, COUNT, REFILL_MIN, HYST, CHUNK>(sc, io, n, work, base, kBlock * 16u, gcounters);
}

// The dual-slot machine (trace_ds.cuh): one ray per phase per lane; the rays' read-only halves in dynamic shared memory.
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, bool OCT, int REFILL_MIN, int MINB, unsigned CHUNK>
__global__ void __launch_bounds__(kBlock, MINB)
    k_trace_ds(const __grid_constant__ SceneView sc, const __grid_constant__ IO io, unsigned long long n,
               const unsigned int *__restrict__ n_dev, unsigned long long *__restrict__ work,
               unsigned long long *__restrict__ gcounters) {
  extern __shared__ uint4 smem_fat[]; // [2 rays x kFatUnits][kBlock], column layout
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_fat + threadIdx.x);
  if (n_dev) n = __ldg(n_dev);
  trace_dual_slot_machine<IO, TRI, CAP, ANYHIT, COUNT, OCT, REFILL_MIN, CHUNK>(sc, io, n, work, base, kBlock * 16u, gcounters);
}

// The phase-sorted machine (trace_ps.cuh): rays live in shared-memory slots and move between the warps of the CTA.
template <class IO, int TRI, int NS, int S, bool ANYHIT, int STAY, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
    k_trace_ps(const __grid_constant__ SceneView sc, const __grid_constant__ IO io, unsigned long long n,
               const unsigned int *__restrict__ n_dev, unsigned long long *__restrict__ work, uint4 *__restrict__ ovf) {
  extern __shared__ __align__(16) unsigned char smem_ps[];
  if (n_dev) n = __ldg(n_dev);
  trace_phase_sorted<IO, TRI, NS, S, 64, ANYHIT, STAY, kChunk>(sc, io, n, work, smem_ps, ovf + (size_t)blockIdx.x * NS * (64 - S));
}

#endif // MB200_DEV_VARIANTS

// ---------------------------------------------------------------------------
// K3: BuildIntersection for a buffer of rays + hit records -> full 184-byte Intersection records
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_isects(const __grid_constant__ SceneView sc,
                                                      const mb200_ray *__restrict__ rays,
                                                      const mb200_hit *__restrict__ hits, size_t n,
                                                      mb200_isect *__restrict__ isects,
                                                      unsigned char *__restrict__ mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 *hp = reinterpret_cast<const double2 *>(hits + i);
    const double2 h0 = __ldg(hp), h1 = __ldg(hp + 1);
    const unsigned long long ids = (unsigned long long)__double_as_longlong(h1.y);
    const uint32_t face = (uint32_t)ids, mat = (uint32_t)(ids >> 32);
    const bool hit = face != 0xFFFFFFFFu;
    mb200_isect o;
    memset(&o, 0, sizeof(o));
    o.t = h0.x, o.u = h0.y, o.v = h1.x, o.faceID = face, o.materialID = mat;
    if (hit) {
      const double2 *rp = reinterpret_cast<const double2 *>(rays + i);
      const double2 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      IsectD d;
      build_intersection(sc, a.x, a.y, b.x, b.y, c.x, c.y, h0.x, h0.y, h1.x, face, d);
      o.f0 = d.f0, o.f1 = d.f1, o.f2 = d.f2;
      o.position[0] = d.px, o.position[1] = d.py, o.position[2] = d.pz;
      o.geometricNormal[0] = d.gx, o.geometricNormal[1] = d.gy, o.geometricNormal[2] = d.gz;
      o.normal[0] = d.nx, o.normal[1] = d.ny, o.normal[2] = d.nz;
      o.texcoord[0] = d.tu, o.texcoord[1] = d.tv;
    }
    isects[i] = o;
    if (mask) mask[i] = hit ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------
// K5: the frame wavefront.  Work item i = one sample of one pixel (shade.cuh: FrameMap).
//
//   trace<IOCamera>        camera ray generated in the refill step, hits[i] <- 32-byte record
//   k_shade_primary        hit record -> PRIMARY_ONLY: contribution | PRIMARY_SHADOW: shadow ray queued |
//                          PATHTRACE: first bounce queued, PathState saved
//   trace<IOQueueShadow>   unoccluded rays deposit their value into contrib[item]
//   trace<IOQueueClosest> / k_shade_bounce   one pair per further path segment (render.cc:401-450)
//   k_resolve              per pixel: contributions of the batch's passes added in pass order, image/count written
// ---------------------------------------------------------------------------
// Appends a ray to a queue: one atomicAdd per warp.  Every lane of the warp must call it.
__device__ __forceinline__ uint32_t queue_slot(unsigned int *qcount, bool want) {
  const unsigned m = __ballot_sync(kFullMask, want);
  if (!m) return 0u;
  const int leader = __ffs(m) - 1;
  unsigned base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(qcount, (unsigned)__popc(m));
  base = __shfl_sync(kFullMask, base, leader);
  return base + (unsigned)__popc(m & ((1u << lane_id()) - 1u));
}

// The same for a whole CTA (every thread must call it), with the CTA's rays grouped by the octant of their
// direction: slots [base + first[oct] ...) in arrival order.  A warp of the next traversal launch then pulls rays
// that walk the same octant copy of the pair nodes, at no extra pass over the queue.
__device__ __forceinline__ uint32_t queue_slot_by_octant(unsigned int *qcount, bool want, double dx, double dy, double dz) {
  __shared__ unsigned int s_cnt[8], s_first[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t oct = (dx < 0.0 ? 1u : 0u) | (dy < 0.0 ? 2u : 0u) | (dz < 0.0 ? 4u : 0u);
  unsigned int rank = 0;
  if (want) rank = atomicAdd(&s_cnt[oct], 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int first[8], total = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) first[k] = total, total += s_cnt[k];
    const unsigned int base = total ? atomicAdd(qcount, total) : 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) s_first[k] = base + first[k];
  }
  __syncthreads();
  return s_first[oct] + rank;
}

__device__ __forceinline__ void store_qray(QRay *dst, double ox, double oy, double oz, double dx, double dy, double dz,
                                           double tmax, uint32_t item, float value) {
  double2 *o = reinterpret_cast<double2 *>(dst);
  o[0] = make_double2(ox, oy);
  o[1] = make_double2(oz, dx);
  o[2] = make_double2(dy, dz);
  o[3] = make_double2(tmax, __longlong_as_double((long long)(((unsigned long long)__float_as_uint(value) << 32) | item)));
}

__device__ __forceinline__ void warp_add_stats(unsigned long long *g, unsigned int a, unsigned int b) {
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(kFullMask, a, o);
    b += __shfl_down_sync(kFullMask, b, o);
  }
  if (lane_id() == 0) {
    if (a) atomicAdd(&g[0], (unsigned long long)a);
    if (b) atomicAdd(&g[3], (unsigned long long)b);
  }
}

// The tail of PathTrace after the path escaped at segment `len` (render.cc:407-417 has no break): the
// remaining segments cannot hit and add throughput*0.5/pathLength each.  Returns the
// number of those "zombie" segments (not rays).
__device__ __forceinline__ unsigned int zombie_tail(unsigned int len, unsigned int max_len, uint32_t cur_mat,
                                                    double &thr, double &radiance) {
  unsigned int z = 0;
  for (;;) {
    if (len >= max_len) return z;
    if (cur_mat != 0xFFFFFFFFu) thr *= 0.5;
    ++len;
    z++;
    radiance += thr * 0.5 / (double)len;
  }
}

// Continuation of a path whose segment `len` hit at (hx,hy,hz) with shading normal n (render.cc:423-449).
__device__ __forceinline__ void bounce_ray(Xorshift128 &rng, double dx, double dy, double dz, double hx, double hy,
                                           double hz, double nx, double ny, double nz, double &ox, double &oy,
                                           double &oz, double &sx, double &sy, double &sz) {
  (void)rng.next(); // drawn, unused (render.cc:430)
  if ((nx * (-dx) + ny * (-dy) + nz * (-dz)) < 0.0) nx = -nx, ny = -ny, nz = -nz;
  sample_diffuse(rng, nx, ny, nz, sx, sy, sz);
  ox = hx + sx * kRenderEPS, oy = hy + sy * kRenderEPS, oz = hz + sz * kRenderEPS;
}

__device__ __forceinline__ void load_hit(const mb200_hit *src, double &t, double &u, double &v, uint32_t &face,
                                         uint32_t &mat) {
  const double2 *hp = reinterpret_cast<const double2 *>(src);
  const double2 h0 = __ldg(hp), h1 = __ldg(hp + 1);
  const unsigned long long ids = (unsigned long long)__double_as_longlong(h1.y);
  t = h0.x, u = h0.y, v = h1.x, face = (uint32_t)ids, mat = (uint32_t)(ids >> 32);
}

__device__ __forceinline__ void save_state(PathState *dst, const Xorshift128 &rng, double thr, double radiance,
                                           uint32_t cur_mat) {
  uint4 *o = reinterpret_cast<uint4 *>(dst);
  o[0] = make_uint4(rng.x, rng.y, rng.z, rng.w);
  reinterpret_cast<double2 *>(dst)[1] = make_double2(thr, radiance);
  o[2] = make_uint4(cur_mat, 0u, 0u, 0u);
}

// One thread per work item, after the camera-ray trace.
__global__ void __launch_bounds__(256)
    k_shade_primary(const __grid_constant__ SceneView sc, const __grid_constant__ mb200_render_params p,
                    const __grid_constant__ FrameMap m, uint32_t items, const mb200_hit *__restrict__ hits,
                    float *__restrict__ contrib, QRay *__restrict__ queue, unsigned int *__restrict__ qcount,
                    PathState *__restrict__ states, unsigned long long *__restrict__ stats, int group) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; // grid covers whole warps of items
  int x = 0, y = 0, rl = 0;
  uint32_t pass = 0;
  const bool valid = i < items && item_pixel(m, i, x, y, rl, pass);
  bool push = false;
  float value = 0.f, out = 0.f;
  double qox = 0, qoy = 0, qoz = 0, qdx = 0, qdy = 0, qdz = 0, qtmax = 0;
  unsigned int zombies = 0;
  if (valid) {
    Xorshift128 rng;
    double ox, oy, oz, dx, dy, dz;
    camera_sample(p, x, y, pass, rng, ox, oy, oz, dx, dy, dz);
    const bool env = p.shader == MB200_SHADER_PATHTRACE_ENV;
    double t, u, v;
    uint32_t face, mat;
    load_hit(hits + i, t, u, v, face, mat);
    bool hit = face != 0xFFFFFFFFu;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    uint32_t cur_mat = 0; // Intersection::materialID before the first Trace (zero-initialised isect)
    if (hit) {
      IsectD d;
      build_intersection(sc, ox, oy, oz, dx, dy, dz, t, u, v, face, d);
      nx = d.nx, ny = d.ny, nz = d.nz;
      cur_mat = mat;
    }
    if (p.use_plane && !env) hit |= plane_intersect(p.plane, ox, oy, oz, dx, dy, dz, t, nx, ny, nz, cur_mat);
    if (env) cur_mat = 0xFFFFFFFFu; // PathTraceEnv never attenuates (render.cc:518-590)
    if (hit) {
      const double hx = ox + t * dx, hy = oy + t * dy, hz = oz + t * dz;
      if (p.shader == MB200_SHADER_PRIMARY_ONLY) {
        out = 1.f;
      } else if (p.shader == MB200_SHADER_PRIMARY_SHADOW) {
        // the next-event-estimation block the reference leaves empty (render.cc:425-426)
        if ((nx * (-dx) + ny * (-dy) + nz * (-dz)) < 0.0) nx = -nx, ny = -ny, nz = -nz;
        double lx = p.light[0] - hx, ly = p.light[1] - hy, lz = p.light[2] - hz;
        const double dist = sqrt(lx * lx + ly * ly + lz * lz);
        normalize3(lx, ly, lz);
        qox = hx + lx * kRenderEPS, qoy = hy + ly * kRenderEPS, qoz = hz + lz * kRenderEPS;
        qdx = lx, qdy = ly, qdz = lz;
        qtmax = dist - kRenderEPS;
        const double ndotl = nx * lx + ny * ly + nz * lz;
        const double kd = (cur_mat != 0xFFFFFFFFu) ? 0.5 : 1.0; // default Material::diffuse (scene.h:58-65)
        value = (ndotl > 0.0) ? (float)(kd * ndotl) : 0.f;
        push = true;
      } else if (p.max_path_length > 1) { // PathTrace, segment 1 hit: continue the path
        double thr = 1.0;
        bounce_ray(rng, dx, dy, dz, hx, hy, hz, nx, ny, nz, qox, qoy, qoz, qdx, qdy, qdz);
        if (cur_mat != 0xFFFFFFFFu) thr *= 0.5;
        qtmax = DBL_MAX;
        save_state(states + i, rng, thr, 0.0, cur_mat);
        push = true;
      }
    }
    // a camera ray that escapes contributes nothing (pathLength < kMinPathLength, render.cc:409-412)
  }
  const uint32_t slot = group ? queue_slot_by_octant(qcount, push, qdx, qdy, qdz) : queue_slot(qcount, push);
  if (push) store_qray(queue + slot, qox, qoy, qoz, qdx, qdy, qdz, qtmax, i, value);
  if (i < items) contrib[i] = out;
  warp_add_stats(stats, valid ? 1u : 0u, zombies);
}

// One thread per traced continuation ray of path segment `len` (>= 2): queue slot j, hits[j].
__global__ void __launch_bounds__(256)
    k_shade_bounce(const __grid_constant__ SceneView sc, const __grid_constant__ mb200_render_params p,
                   unsigned int len, const QRay *__restrict__ qin, const unsigned int *__restrict__ qin_count,
                   const mb200_hit *__restrict__ hits, QRay *__restrict__ qout, unsigned int *__restrict__ qout_count,
                   PathState *__restrict__ states, float *__restrict__ contrib,
                   unsigned long long *__restrict__ stats, int group) {
  const unsigned int n = __ldg(qin_count);
  const unsigned int max_len = (unsigned int)p.max_path_length;
  unsigned int zombies = 0, traced = 0;
  const unsigned int stride = gridDim.x * blockDim.x;
  for (unsigned int base = blockIdx.x * blockDim.x; base < n; base += stride) { // uniform per CTA
    const unsigned int j = base + threadIdx.x;
    bool push = false;
    double qox = 0, qoy = 0, qoz = 0, qdx = 0, qdy = 0, qdz = 0;
    uint32_t item = 0;
    if (j < n) {
      traced++;
      const double2 *qp = reinterpret_cast<const double2 *>(qin + j);
      const double2 a = __ldg(qp), b = __ldg(qp + 1), c = __ldg(qp + 2), d3 = __ldg(qp + 3);
      const double ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
      item = (uint32_t)(unsigned long long)__double_as_longlong(d3.y);
      PathState *sp = states + item;
      const uint4 s0 = *reinterpret_cast<const uint4 *>(sp);
      const double2 s1 = reinterpret_cast<const double2 *>(sp)[1];
      uint32_t cur_mat = reinterpret_cast<const uint4 *>(sp)[2].x;
      Xorshift128 rng{s0.x, s0.y, s0.z, s0.w};
      double thr = s1.x, radiance = s1.y;

      double t, u, v;
      uint32_t face, mat;
      load_hit(hits + j, t, u, v, face, mat);
      bool hit = face != 0xFFFFFFFFu;
      double nx = 0.0, ny = 0.0, nz = 0.0;
      if (hit) {
        IsectD d;
        build_intersection(sc, ox, oy, oz, dx, dy, dz, t, u, v, face, d);
        nx = d.nx, ny = d.ny, nz = d.nz;
        cur_mat = mat;
      }
      const bool env = p.shader == MB200_SHADER_PATHTRACE_ENV;
      if (p.use_plane && !env) hit |= plane_intersect(p.plane, ox, oy, oz, dx, dy, dz, t, nx, ny, nz, cur_mat);
      if (env) cur_mat = 0xFFFFFFFFu;
      if (!hit) { // escaped: this segment and every later one adds throughput*0.5/pathLength
        radiance += thr * 0.5 / (double)len;
        zombies += zombie_tail(len, max_len, cur_mat, thr, radiance);
        contrib[item] = (float)radiance;
      } else if (len >= max_len) {
        contrib[item] = (float)radiance;
      } else {
        const double hx = ox + t * dx, hy = oy + t * dy, hz = oz + t * dz;
        bounce_ray(rng, dx, dy, dz, hx, hy, hz, nx, ny, nz, qox, qoy, qoz, qdx, qdy, qdz);
        if (cur_mat != 0xFFFFFFFFu) thr *= 0.5;
        save_state(sp, rng, thr, radiance, cur_mat);
        push = true;
      }
    }
    const uint32_t slot = group ? queue_slot_by_octant(qout_count, push, qdx, qdy, qdz) : queue_slot(qout_count, push);
    if (push) store_qray(qout + slot, qox, qoy, qoz, qdx, qdy, qdz, DBL_MAX, item, 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    traced += __shfl_down_sync(kFullMask, traced, o);
    zombies += __shfl_down_sync(kFullMask, zombies, o);
  }
  if (lane_id() == 0) {
    if (traced) atomicAdd(&stats[1], (unsigned long long)traced);
    if (zombies) atomicAdd(&stats[3], (unsigned long long)zombies);
  }
}

// ---- bounce-queue re-ordering (SURVEY 7.2 "Divergence"; the reference's bounce loop is render.cc:401-450) --------------
// Continuation rays leave k_shade_* in sample order: neighbouring slots hold rays with unrelated directions.  A single
// counting-sort pass regroups the queue by  key = octant of the direction (3 bits: the octant copy of the pair nodes the
// ray will walk, layout.h) << 12 | Morton code of the origin's cell in a 16^3 grid over the scene box,  so that the 32
// rays a warp pulls share their node copy and start next to each other.  The sample a ray belongs to travels inside
// the QRay, so nothing downstream depends on the slot a ray sits in; the order inside a bucket is whatever the atomics
// give and is not observable (a ray's result does not depend on its neighbours).
constexpr uint32_t kSortBuckets = 8u << 12;
struct SortGrid {
  double lo[3], scale[3]; // cell = (o - lo) * scale, clamped to 0..15
};

__device__ __forceinline__ uint32_t spread4(uint32_t v) { // abcd -> a00b00c00d
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__device__ __forceinline__ uint32_t sort_key(const SortGrid &g, const QRay *q) {
  const double2 *p = reinterpret_cast<const double2 *>(q);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  const double o[3] = {a.x, a.y, b.x};
  uint32_t cell[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double f = (o[k] - g.lo[k]) * g.scale[k];
    cell[k] = f >= 15.0 ? 15u : (f > 0.0 ? (uint32_t)f : 0u); // NaN -> 0
  }
  const uint32_t oct = (b.y < 0.0 ? 1u : 0u) | (c.x < 0.0 ? 2u : 0u) | (c.y < 0.0 ? 4u : 0u);
  return (oct << 12) | spread4(cell[0]) | (spread4(cell[1]) << 1) | (spread4(cell[2]) << 2);
}

__global__ void __launch_bounds__(256) k_sort_count(const __grid_constant__ SortGrid g, const QRay *__restrict__ q,
                                                    const unsigned int *__restrict__ n_dev, unsigned int *__restrict__ hist) {
  const unsigned int n = __ldg(n_dev);
  for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    atomicAdd(&hist[sort_key(g, q + j)], 1u);
}

// exclusive scan of the kSortBuckets counters in place: one CTA, 32 counters per thread
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned int *__restrict__ hist) {
  __shared__ unsigned int warp_sums[32];
  constexpr unsigned int kPer = kSortBuckets / 1024;
  const unsigned int t = threadIdx.x, lane = t & 31u, w = t >> 5;
  unsigned int v[kPer], sum = 0;
#pragma unroll
  for (unsigned int k = 0; k < kPer; k += 4) {
    const uint4 x = reinterpret_cast<const uint4 *>(hist + t * kPer)[k / 4];
    v[k] = x.x, v[k + 1] = x.y, v[k + 2] = x.z, v[k + 3] = x.w;
    sum += x.x + x.y + x.z + x.w;
  }
  unsigned int inc = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int up = __shfl_up_sync(kFullMask, inc, o);
    if (lane >= (unsigned)o) inc += up;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned int ws = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(kFullMask, ws, o);
      if (lane >= (unsigned)o) ws += up;
    }
    warp_sums[lane] = ws;
  }
  __syncthreads();
  unsigned int run = inc - sum + (w ? warp_sums[w - 1] : 0u);
#pragma unroll
  for (unsigned int k = 0; k < kPer; k++) {
    const unsigned int c = v[k];
    hist[t * kPer + k] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ SortGrid g, const QRay *__restrict__ q,
                                                      const unsigned int *__restrict__ n_dev, unsigned int *__restrict__ offs,
                                                      QRay *__restrict__ out) {
  const unsigned int n = __ldg(n_dev);
  for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const unsigned int pos = atomicAdd(&offs[sort_key(g, q + j)], 1u);
    const uint4 *src = reinterpret_cast<const uint4 *>(q + j);
    uint4 *dst = reinterpret_cast<uint4 *>(out + pos);
    const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
    dst[0] = a, dst[1] = b, dst[2] = c, dst[3] = d;
  }
}

// One thread per (tile, lane) pixel: the batch's samples are added in pass order, float += float
// (AccumImage, main_sdl.cc:138-143), then image / count are written as `mode` says (kernels.h).
__global__ void __launch_bounds__(256)
    k_resolve(const __grid_constant__ FrameMap m, uint32_t tiles, int mode, const float *__restrict__ contrib,
              float *__restrict__ image, int *__restrict__ count) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t tile = g >> 5, lane = g & 31u;
  if (tile >= tiles) return;
  int x, y, rl;
  uint32_t pass;
  if (!item_pixel(m, (tile * m.passes) * 32u + lane, x, y, rl, pass)) return;
  const size_t pix = pixel_slot(m, x, y, rl);
  float ar = 0.f, ag = 0.f, ab = 0.f;
  if (mode == 1) ar = image[3 * pix + 0], ag = image[3 * pix + 1], ab = image[3 * pix + 2];
  const float *c = contrib + (size_t)tile * m.passes * 32u + lane;
  for (uint32_t s = 0; s < m.passes; s++) {
    const float r = c[(size_t)s * 32u];
    if (mode != 0) ar += r, ag += r, ab += r;
    else ar = ag = ab = r;
  }
  if (m.step > 1) {
    // coarse preview (render.cc:684-696): the sample fills its step x step block, clipped to the
    // rectangle; count++ sits inside the reference's k < 3 colour loop, hence += 3 per pass.
    const int bx = min(m.step, m.x1 - x), by = min(m.step, m.y1 - y);
    for (int v = 0; v < by; v++)
      for (int u = 0; u < bx; u++) {
        const size_t q = (size_t)(y + v) * m.width + (x + u);
        image[3 * q + 0] = ar, image[3 * q + 1] = ag, image[3 * q + 2] = ab;
        if (mode == 2) count[q] = 3 * (int)m.passes;
        else count[q] += 3 * (int)m.passes;
      }
    return;
  }
  image[3 * pix + 0] = ar, image[3 * pix + 1] = ag, image[3 * pix + 2] = ab;
  if (mode == 2) count[pix] = (int)m.passes;
  else count[pix] += (int)m.passes;
}

// ---------------------------------------------------------------------------
// Output resolve: accumulated float frame + per-pixel sample counts -> 8-bit pixels, as the reference's two
// front ends do it on the host.
//   mode 0  HDRToLDR (main_console.cc:25-43): RGB8, fclamp(in / count) with int i = x * 255.5 (float * double)
//   mode 1  Display  (main_sdl.cc:156-165, 420-477): BGRA8 (B at byte 0, A = 255), scale = 1.0f / (float)count,
//           fclamp(scale * in) with int i = powf(x, 1.0f / 2.2f) * 255.5
// `(int)double` is x86-64's cvttsd2si in the reference build: NaN and out-of-range values give INT_MIN, which the
// clamp turns into 0 (CUDA's own conversion would saturate to 255 / 0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int quantise_x86(double v) {
  const int i = (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
  return i < 0 ? 0u : (i > 255 ? 255u : (unsigned int)i);
}

// powf(x, 1.0f / 2.2f) evaluated as pow in double, rounded to float once: equals glibc's powf (itself computed in
// double) except where the two double results straddle a float rounding boundary.
__device__ __forceinline__ float gamma22(float x) { return (float)pow((double)x, (double)(1.0f / 2.2f)); }

__global__ void __launch_bounds__(256) k_resolve_ldr(const float *__restrict__ image, const int *__restrict__ count,
                                                     size_t npix, int mode, unsigned char *__restrict__ out) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    const float r = image[3 * p + 0], g = image[3 * p + 1], b = image[3 * p + 2];
    const int n = count[p];
    if (mode == 0) {
      const float fn = (float)n; // in[i] / in_count[i / 3]: float / int
      out[3 * p + 0] = (unsigned char)quantise_x86((double)(r / fn) * 255.5);
      out[3 * p + 1] = (unsigned char)quantise_x86((double)(g / fn) * 255.5);
      out[3 * p + 2] = (unsigned char)quantise_x86((double)(b / fn) * 255.5);
    } else {
      const float scale = 1.0f / (float)n;
      const unsigned int qr = quantise_x86((double)gamma22(scale * r) * 255.5);
      const unsigned int qg = quantise_x86((double)gamma22(scale * g) * 255.5);
      const unsigned int qb = quantise_x86((double)gamma22(scale * b) * 255.5);
      reinterpret_cast<unsigned int *>(out)[p] = qb | (qg << 8) | (qr << 16) | 0xFF000000u;
    }
  }
}

// order[tile0 .. tile0 + tiles) = stable partition of the tiles tile0 .. tile0 + tiles - 1: those flagged in `hot` first,
// then the others; clears their flags.  One CTA of 1024 threads per range (a 1080p frame has 64 800 tiles: 64 rounds).
struct OrderRanges { // one CTA per batch of the frame
  uint32_t tile0[64], tiles[64];
};
__global__ void __launch_bounds__(1024) k_build_order(unsigned char *__restrict__ hot, const __grid_constant__ OrderRanges rg,
                                                      uint32_t *__restrict__ order) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t base_hot, base_cold;
  const uint32_t tile0 = rg.tile0[blockIdx.x], tiles = rg.tiles[blockIdx.x];
  hot += tile0, order += tile0;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  // pass 1: how many hot tiles
  uint32_t mine = 0;
  for (uint32_t i = tid; i < tiles; i += 1024u) mine += hot[i] ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(kFullMask, mine, o);
  if (lane == 0) warp_sums[warp] = mine;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 32; w++) t += warp_sums[w];
    base_hot = 0, base_cold = t;
  }
  __syncthreads();
  // pass 2: rounds of 1024 tiles, each a block-wide exclusive scan of the hot flags
  for (uint32_t start = 0; start < tiles; start += 1024u) {
    const uint32_t i = start + tid;
    const bool in = i < tiles;
    const bool h = in && hot[i] != 0;
    const unsigned bal = __ballot_sync(kFullMask, h);
    const uint32_t before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    uint32_t warp_off = 0, round_hot = 0;
    for (uint32_t w = 0; w < 32u; w++) {
      const uint32_t c = warp_sums[w];
      if (w < warp) warp_off += c;
      round_hot += c;
    }
    const uint32_t hot_rank = warp_off + before;   // hot tiles of this round before me
    const uint32_t cold_rank = tid - hot_rank;     // cold tiles of this round before me
    if (in) order[h ? base_hot + hot_rank : base_cold + cold_rank] = tile0 + i;
    if (in) hot[i] = 0;
    __syncthreads();
    if (tid == 0) {
      const uint32_t n = (tiles - start < 1024u) ? tiles - start : 1024u;
      base_hot += round_hot;
      base_cold += n - round_hot;
    }
    __syncthreads();
  }
}

// tcount: the traversal launches' counters, camera [0..3] and shadow [4..7] (nodes, tris, rays, max stack)
__global__ void k_add_stats(const unsigned long long *__restrict__ batch, const unsigned int *__restrict__ shadow_count,
                            const unsigned long long *__restrict__ tcount, unsigned long long *__restrict__ total) {
  // batches of a frame run on two streams: accumulate atomically
  if (threadIdx.x < 4 && batch[threadIdx.x]) atomicAdd(&total[threadIdx.x], batch[threadIdx.x]);
  if (threadIdx.x == 2 && shadow_count && *shadow_count) atomicAdd(&total[2], (unsigned long long)*shadow_count);
  const int map[4] = {0, 1, 4, 5}; // -> total[4..7] = camera nodes, camera tris, shadow nodes, shadow tris
  if (threadIdx.x >= 4 && threadIdx.x < 8 && tcount[map[threadIdx.x - 4]]) atomicAdd(&total[threadIdx.x], tcount[map[threadIdx.x - 4]]);
}

// fused batch: counters of the state machine (trace_sm.cuh: [0..3] camera nodes, tris, rays, max stack; [4..6] shadow
// nodes, tris, rays) -> the caller's stats [primary, bounce, shadow, zombie, camera nodes, camera tris, shadow
// nodes, shadow tris]
__global__ void k_add_stats_fused(const unsigned long long *__restrict__ batch, unsigned long long *__restrict__ total) {
  const int map[8] = {2, -1, 6, -1, 0, 1, 4, 5};
  const int t = threadIdx.x;
  if (t < 8 && map[t] >= 0 && batch[map[t]]) atomicAdd(&total[t], batch[map[t]]);
}

// Octant copies of the pair nodes (layout.h: nodes_oct): one thread per (node, octant).
__global__ void __launch_bounds__(256) k_octant_nodes(const PairNode *__restrict__ src, uint32_t n, PairNode *__restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * 8) return;
  const uint32_t oct = (uint32_t)(i / n), k = (uint32_t)(i - (size_t)oct * n);
  const PairNode s = src[k];
  const bool swap = ((oct >> s.axis) & 1u) != 0u; // dirSign[axis] = 1: data[1] is the near child (bvh_accel.cc:818-823)
  PairNode o;
  for (int c = 0; c < 2; c++) {
    const int from = swap ? 1 - c : c;
    for (int a = 0; a < 3; a++) {
      const bool neg = ((oct >> a) & 1u) != 0u; // IntersectRayAABB: min plane = dirSign ? bmax : bmin (bvh_accel.cc:556-561)
      o.box[c][a] = neg ? s.box[from][3 + a] : s.box[from][a];
      o.box[c][3 + a] = neg ? s.box[from][a] : s.box[from][3 + a];
    }
    // a branch child is addressed inside the same copy: absolute index, so the walk needs no per-ray offset
    o.ref[c] = s.cnt[from] == kBranch ? oct * n + s.ref[from] : s.ref[from];
    o.cnt[c] = s.cnt[from];
  }
  o.axis = s.axis;
  o.pad_[0] = o.pad_[1] = o.pad_[2] = 0u;
  dst[i] = o;
}

// 64-byte copies of the pair nodes (layout.h: PairNode64).  *bad is set when a box coordinate is not (float -/+ kEPS)
// exactly or a leaf holds more than 65 534 triangles: the scene then keeps its 128-byte nodes.
__global__ void __launch_bounds__(256) k_pack_nodes64(const PairNode *__restrict__ src, uint32_t n, PairNode64 *__restrict__ dst,
                                                      int *__restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PairNode s = src[i];
  PairNode64 o;
  bool ok = true;
  for (int c = 0; c < 2; c++) {
    for (int k = 0; k < 6; k++) {
      const double v = s.box[c][k];
      const float f = (float)(k < 3 ? v + MB200_TRI_EPS : v - MB200_TRI_EPS);
      const double back = k < 3 ? (double)f - MB200_TRI_EPS : (double)f + MB200_TRI_EPS;
      ok &= __double_as_longlong(back) == __double_as_longlong(v);
      o.box[c][k] = f;
    }
    o.ref[c] = s.ref[c];
    ok &= s.cnt[c] == kBranch || s.cnt[c] < 0xFFFFu;
    o.cnt[c] = s.cnt[c] == kBranch ? (uint16_t)0xFFFFu : (uint16_t)s.cnt[c];
  }
  o.axis = s.axis;
  dst[i] = o;
  if (!ok) atomicOr(bad, 1);
}

// Traversal copies of the triangle records (layout.h: TriKind).  One thread per record.
__global__ void __launch_bounds__(256) k_pad_tris(const void *__restrict__ src, int src_f32, uint32_t n, int kind,
                                                  void *__restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (kind == kTriF32x64) {
    const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const TriRecordF32 *>(src) + i);
    uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(dst) + (size_t)i * 64u);
    o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  double v[12];
  uint32_t face, mat;
  if (src_f32) {
    const TriRecordF32 t = reinterpret_cast<const TriRecordF32 *>(src)[i];
    for (int c = 0; c < 3; c++) {
      v[c] = (double)t.p0[c];
      v[3 + c] = (double)t.p1[c] - (double)t.p0[c]; // TriangleIsect's e1, e2 (bvh_accel.cc:600-603)
      v[6 + c] = (double)t.p2[c] - (double)t.p0[c];
    }
    face = t.face, mat = t.mat;
  } else {
    const TriRecordF64 t = reinterpret_cast<const TriRecordF64 *>(src)[i];
    for (int c = 0; c < 3; c++) v[c] = t.p0[c], v[3 + c] = t.e1[c], v[6 + c] = t.e2[c];
    face = t.face, mat = t.mat;
  }
  double *o = reinterpret_cast<double *>(reinterpret_cast<char *>(dst) + (size_t)i * 96u);
  if (kind == kTriWoop) { // traverse.cuh: rows r1, r2, r3 = n of the map into the unit triangle, offsets b = -r . p0
    const double *p0 = v, *e1 = v + 3, *e2 = v + 6;
    const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double a[3] = {e2[1] * n[2] - e2[2] * n[1], e2[2] * n[0] - e2[0] * n[2], e2[0] * n[1] - e2[1] * n[0]}; // e2 x n
    const double b[3] = {n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2], n[0] * e1[1] - n[1] * e1[0]}; // n x e1
    const double sa = 1.0 / ((e1[0] * a[0] + e1[1] * a[1]) + e1[2] * a[2]);
    const double sb = 1.0 / ((e2[0] * b[0] + e2[1] * b[1]) + e2[2] * b[2]);
    double w[12];
    for (int c = 0; c < 3; c++) w[c] = a[c] * sa, w[4 + c] = b[c] * sb, w[8 + c] = n[c];
    w[3] = -((w[0] * p0[0] + w[1] * p0[1]) + w[2] * p0[2]);
    w[7] = -((w[4] * p0[0] + w[5] * p0[1]) + w[6] * p0[2]);
    w[11] = -((w[8] * p0[0] + w[9] * p0[1]) + w[10] * p0[2]);
    for (int c = 0; c < 12; c++) o[c] = w[c];
    return;
  }
  v[9] = __longlong_as_double((long long)(((unsigned long long)mat << 32) | face));
  v[10] = v[11] = 0.0;
  for (int c = 0; c < 12; c++) o[c] = v[c];
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
std::atomic<int> g_num_sms[64]; // per device ordinal; hosts may drive one scene per thread
std::atomic<int> g_launches{0};

int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = (dev >= 0 && dev < 64) ? dev : 0;
  int n = g_num_sms[slot].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_num_sms[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Function attributes (dynamic shared memory opt-in, carve-out) are per device: a process that drives several
// GPUs (mb200_render_frame_multi) sets them once on each.  Returns the persistent grid for the current device.
template <class Kernel>
cudaError_t persistent_grid(Kernel k, size_t smem, int minb, int grids[64], int *grid_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (grids[dev] == 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // shared memory for `minb` resident CTAs and no more: what is left of the SM's 256 KB stays L1, which
    // is what the node / triangle fetches hit in
    int carve = (int)((smem + 1024) * minb * 100 / (228 * 1024)) + 1;
    if (carve > 100) carve = 100;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kBlock, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > minb) per_sm = minb;
    grids[dev] = (sms > 0 ? sms : 148) * per_sm;
  }
  *grid_out = grids[dev];
  return cudaSuccess;
}

// One persistent wave: a whole number of CTAs per SM.
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int SHADE_MIN, int S, int MINB, unsigned CHUNK, int VAR>
cudaError_t launch_sm(const SceneView &sc, const IO &io, size_t n, const unsigned int *n_dev,
                      unsigned long long *work, unsigned long long *counters, cudaStream_t s) {
  auto k = k_trace_sm<IO, TRI, CAP, ANYHIT, COUNT, REFILL_MIN, SHADE_MIN, S, MINB, CHUNK, VAR>;
  size_t smem = (size_t)(S + (IO::kFused ? 2 : 0)) * kBlock * sizeof(uint4);
  if (VAR & kVarTopSmem) smem += 16 + (size_t)sc.top_count * kTopSlotBytes;
  static int grids[64]; // per instantiation, per device
  int grid = 0;
  const cudaError_t ge = persistent_grid(k, smem, MINB, grids, &grid);
  if (ge != cudaSuccess) return ge;
  k<<<grid, kBlock, smem, s>>>(sc, io, (unsigned long long)n, n_dev, work, counters);
  g_launches++;
  return cudaGetLastError();
}

#ifdef MB200_DEV_VARIANTS
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int HYST, int MINB, unsigned CHUNK>
cudaError_t launch_tr(const SceneView &sc, const IO &io, size_t n, const unsigned int *n_dev, unsigned long long *work,
                      unsigned long long *counters, cudaStream_t s) {
  auto k = k_trace_tr<IO, TRI, CAP, ANYHIT, COUNT, REFILL_MIN, HYST, MINB, CHUNK>;
  const size_t smem = (size_t)2 * kFatUnits * kBlock * sizeof(uint4);
  static int grids[64]; // per instantiation, per device
  int grid = 0;
  const cudaError_t ge = persistent_grid(k, smem, MINB, grids, &grid);
  if (ge != cudaSuccess) return ge;
  k<<<grid, kBlock, smem, s>>>(sc, io, (unsigned long long)n, n_dev, work, counters);
  g_launches++;
  return cudaGetLastError();
}

template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, bool OCT, int REFILL_MIN, int MINB, unsigned CHUNK>
cudaError_t launch_ds(const SceneView &sc, const IO &io, size_t n, const unsigned int *n_dev, unsigned long long *work,
                      unsigned long long *counters, cudaStream_t s) {
  auto k = k_trace_ds<IO, TRI, CAP, ANYHIT, COUNT, OCT, REFILL_MIN, MINB, CHUNK>;
  const size_t smem = (size_t)2 * kFatUnits * kBlock * sizeof(uint4);
  static int grids[64]; // per instantiation, per device
  int grid = 0;
  const cudaError_t ge = persistent_grid(k, smem, MINB, grids, &grid);
  if (ge != cudaSuccess) return ge;
  k<<<grid, kBlock, smem, s>>>(sc, io, (unsigned long long)n, n_dev, work, counters);
  g_launches++;
  return cudaGetLastError();
}

template <class IO, int TRI, int NS, int S, bool ANYHIT, int STAY, int MINB>
cudaError_t launch_ps(const SceneView &sc, const IO &io, size_t n, const unsigned int *n_dev, unsigned long long *work,
                      cudaStream_t s) {
  auto k = k_trace_ps<IO, TRI, NS, S, ANYHIT, STAY, MINB>;
  const size_t smem = ((sizeof(PsShared) + 3 * NS * 2 + 15) & ~(size_t)15) + kPsStageBytes + (size_t)(kPsRoUnits + kPsMutUnits + S) * NS * 16;
  static int grids[64]; // per instantiation, per device
  int grid = 0;
  const cudaError_t ge = persistent_grid(k, smem, MINB, grids, &grid);
  if (ge != cudaSuccess) return ge;
  // slot-indexed overflow of the traversal stacks (development variant: one allocation per stream that launches this
  // instantiation -- launches of the frame pipeline's two streams overlap -- never freed)
  struct Ovf {
    cudaStream_t stream;
    int device;
    uint4 *p;
    size_t ctas;
  };
  static Ovf table[16];
  static int used = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  Ovf *o = nullptr;
  for (int i = 0; i < used; i++)
    if (table[i].stream == s && table[i].device == dev) o = &table[i];
  if (!o) {
    if (used == 16) return cudaErrorMemoryAllocation;
    o = &table[used++];
    o->stream = s, o->device = dev, o->p = nullptr, o->ctas = 0;
  }
  if (o->ctas < (size_t)grid) {
    if (o->p) cudaFree(o->p);
    o->p = nullptr, o->ctas = 0;
    const cudaError_t e = cudaMalloc((void **)&o->p, (size_t)grid * NS * (64 - S) * sizeof(uint4));
    if (e != cudaSuccess) return e;
    o->ctas = (size_t)grid;
  }
  k<<<grid, kBlock, smem, s>>>(sc, io, (unsigned long long)n, n_dev, work, o->p);
  g_launches++;
  return cudaGetLastError();
}

#endif // MB200_DEV_VARIANTS

// Picks the kernel instantiation: production parameters, or (development builds, -DMB200_DEV_VARIANTS)
// the A/B variants selected with MB200_TRACE_VAR.
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT>
cudaError_t launch_sm_variant(const SceneView &sc, const IO &io, size_t n, const unsigned int *n_dev,
                              unsigned long long *work, unsigned long long *counters, cudaStream_t s) {
  // frames keep 12 stack entries per thread in shared memory; a lone query launch (no second stream to hide its tail
  // behind) measured 7 % faster with 8, i.e. with more of the SM's 256 KB left as L1 (profiles/r2_ab7_*.log)
  constexpr int S = IO::kFused ? kSmemStackFused : (IO::kTracksCost ? kSmemStack : kSmemStackQuery);
#define MB200_SM(R, H, SS, B, C, V) launch_sm<IO, TRI, CAP, ANYHIT, COUNT, R, H, SS, B, C, V>(sc, io, n, n_dev, work, counters, s)
#ifdef MB200_DEV_VARIANTS
  if constexpr (!IO::kFused && CAP <= 64 && (TRI == kTriF32 || TRI == kTriF64)) {
    // two rays per lane, one body per iteration (trace_tr.cuh): MB200_TRACE_TR = hysteresis of the vote (0, 4, 8) + 1
    static const int tr = env_int("MB200_TRACE_TR", 0);
#define MB200_TR(R, H, B) launch_tr<IO, TRI, CAP, ANYHIT, COUNT, R, H, B, kChunk>(sc, io, n, n_dev, work, counters, s)
    // one slot per phase per lane (trace_ds.cuh): MB200_TRACE_DS = CTAs per SM * 100 + refill threshold
    // rays move between the warps of a CTA, one step body per warp round (trace_ps.cuh): MB200_TRACE_PS = slots per CTA
    // * 100 + lanes below which a round ends (19216: 192 slots, 16 lanes)
    static const int ps = env_int("MB200_TRACE_PS", 0);
    if (ps && sc.nodes_oct && !COUNT && TRI == kTriF32) {
      if constexpr (!COUNT && TRI == kTriF32) {
        if (ps == 19216) return launch_ps<IO, TRI, 192, 4, ANYHIT, 16, 6>(sc, io, n, n_dev, work, s);
        if (ps == 19208) return launch_ps<IO, TRI, 192, 4, ANYHIT, 8, 6>(sc, io, n, n_dev, work, s);
        if (ps == 19224) return launch_ps<IO, TRI, 192, 4, ANYHIT, 24, 6>(sc, io, n, n_dev, work, s);
        if (ps == 25616) return launch_ps<IO, TRI, 256, 4, ANYHIT, 16, 5>(sc, io, n, n_dev, work, s);
        if (ps == 16016) return launch_ps<IO, TRI, 160, 4, ANYHIT, 16, 7>(sc, io, n, n_dev, work, s);
      }
    }
    static const int ds = env_int("MB200_TRACE_DS", 0);
    if (ds && sc.nodes_oct) {
#define MB200_DS(R, B) launch_ds<IO, TRI, CAP, ANYHIT, COUNT, true, R, B, kChunk>(sc, io, n, n_dev, work, counters, s)
      if (ds == 608) return MB200_DS(8, 6);
      if (ds == 604) return MB200_DS(4, 6);
      if (ds == 616) return MB200_DS(16, 6);
      if (ds == 508) return MB200_DS(8, 5);
      if (ds == 708) return MB200_DS(8, 7);
      if (ds == 808) return MB200_DS(8, 8);
#undef MB200_DS
    }
    if (tr == 1) return MB200_TR(kRefillMin, 0, kMinBlocks);
    if (tr == 5) return MB200_TR(kRefillMin, 4, kMinBlocks);
    if (tr == 9) return MB200_TR(kRefillMin, 8, kMinBlocks);
    if (tr == 105) return MB200_TR(4, 4, kMinBlocks);
    if (tr == 205) return MB200_TR(16, 4, kMinBlocks);
#undef MB200_TR
  }
  if constexpr (!COUNT && CAP <= 64) {
    static const int var = env_int("MB200_TRACE_VAR", -1); // -1 = not set: the production instantiation
    if (var == 1) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVarOctant);
    if (var == 104 && sc.nodes_oct) return MB200_SM(kRefillMin, 4, S, kMinBlocks, kChunk, kVar);      // shade step at 4 parked lanes
    if (var == 116 && sc.nodes_oct) return MB200_SM(kRefillMin, 16, S, kMinBlocks, kChunk, kVar);     // ... at 16
    if (var == 212 && sc.nodes_oct) return MB200_SM(12, 12, S, kMinBlocks, kChunk, kVar);             // refill and shade at 12
    if (var == 24 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVar | kVarVote);  // phase vote
    if (var == 56 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVar | kVarVote | kVarVoteBoth);
    if (var == 0) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, 0);   // canonical nodes (round-1 production)
    if (var == 4 && sc.nodes64) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVarNode64);  // 64-byte pair nodes
    if (var == 2) return MB200_SM(kRefillMin, kShadeMin, 0, kMinBlocks, kChunk, kVarTopSmem);  // top of the tree in shared memory
    if (var == 1300) return MB200_SM(kRefillMin, kShadeMin, 0, kMinBlocks, kChunk, 0);   // control for it: canonical nodes, stack in local memory
    if (var == 308 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 8, kMinBlocks, kChunk, kVar);  // 8 stack entries in shared memory
    if (var == 316 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 16, kMinBlocks, kChunk, kVar); // 16
    if (var == 300 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 0, kMinBlocks, kChunk, kVar);  // none: stack in local memory
    if (var == 409 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 12, 9, kChunk, kVar);   // 9 CTAs per SM (56 registers)
    if (var == 410 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 10, 10, kChunk, kVar);  // 10 CTAs per SM (48 registers)
    if (var == 407 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 12, 7, kChunk, kVar);   // 7 CTAs per SM (72 registers)
    if (var == 406 && sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, 12, 6, kChunk, kVar);   // 6 CTAs per SM (80 registers)
    if (var == 506 && sc.nodes_oct) return MB200_SM(6, kShadeMin, 12, kMinBlocks, kChunk, kVar);   // refill at 6 idle lanes
    if (var == 512 && sc.nodes_oct) return MB200_SM(12, kShadeMin, 12, kMinBlocks, kChunk, kVar);  // refill at 12
  }
#endif
  // scenes without octant copies (MB200_NODE_OCT=0, or not enough memory for them) walk the canonical nodes
  if (!sc.nodes_oct) return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVar & ~kVarOctNodes);
  return MB200_SM(kRefillMin, kShadeMin, S, kMinBlocks, kChunk, kVar);
#undef MB200_SM
}

// dispatch on the triangle records the traversal reads and on the stack capacity the tree needs.  The padded
// record kinds exist for the frame kernels and the plain queries alike; counting launches (diagnostics) use them too.
template <class IO, bool ANYHIT, bool COUNT>
cudaError_t launch_trace_kind(const SceneView &sc, int stack_cap, const IO &io, size_t n, const unsigned int *n_dev,
                              unsigned long long *work, unsigned long long *counters, cudaStream_t s) {
#define MB200_KIND(K)                                                                                        \
  case K:                                                                                                    \
    return stack_cap <= 64 ? launch_sm_variant<IO, K, 64, ANYHIT, COUNT>(sc, io, n, n_dev, work, counters, s) \
                           : launch_sm_variant<IO, K, 512, ANYHIT, COUNT>(sc, io, n, n_dev, work, counters, s);
  switch (sc.tri_kind) {
    MB200_KIND(kTriF32)
    MB200_KIND(kTriF64)
#ifdef MB200_DEV_VARIANTS // A/B: padded traversal records (MB200_TRI_LAYOUT, scene.cc)
    MB200_KIND(kTriF32x64)
    MB200_KIND(kTriF64x96)
    MB200_KIND(kTriWoop)
#endif
  }
#undef MB200_KIND
  return cudaErrorInvalidValue;
}

template <class IO, bool ANYHIT>
cudaError_t launch_trace(const SceneView &sc, int stack_cap, const IO &io, size_t n, const unsigned int *n_dev,
                         unsigned long long *work, unsigned long long *counters, cudaStream_t s) {
  if (counters) return launch_trace_kind<IO, ANYHIT, true>(sc, stack_cap, io, n, n_dev, work, counters, s);
  return launch_trace_kind<IO, ANYHIT, false>(sc, stack_cap, io, n, n_dev, work, nullptr, s);
}

// frame-internal traces of the wavefront path never count
template <class IO, bool ANYHIT>
cudaError_t launch_trace_nocount(const SceneView &sc, int stack_cap, const IO &io, size_t n, const unsigned int *n_dev,
                                 unsigned long long *work, cudaStream_t s) {
  return launch_trace_kind<IO, ANYHIT, false>(sc, stack_cap, io, n, n_dev, work, nullptr, s);
}

int flat_grid(size_t n, int block) {
  const size_t want = (n + block - 1) / block, cap = (size_t)num_sms() * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace

int launches_issued() { return g_launches.load(); }
void note_launch() { g_launches++; }

// ---------------------------------------------------------------------------
// KernelTimer
// ---------------------------------------------------------------------------
void KernelTimer::begin(int cls, cudaStream_t s) {
  if (!enabled) return;
  Span sp;
  sp.cls = cls;
  for (cudaEvent_t *e : {&sp.a, &sp.b}) {
    if (!pool.empty()) {
      *e = pool.back();
      pool.pop_back();
    } else if (cudaEventCreate(e) != cudaSuccess) {
      enabled = false;
      return;
    }
  }
  cudaEventRecord(sp.a, s);
  spans.push_back(sp);
}

void KernelTimer::end(cudaStream_t s) {
  if (!enabled || spans.empty()) return;
  cudaEventRecord(spans.back().b, s);
}

void KernelTimer::collect(double ms[kKClasses], unsigned long long launches[kKClasses], double *trace_union_ms) {
  if (trace_union_ms) *trace_union_ms = 0.0;
  if (!spans.empty()) {
    // absolute intervals relative to the first recorded event, then the union length per class
    struct Iv {
      float a, b;
      int cls;
    };
    std::vector<Iv> iv;
    const cudaEvent_t ref = spans[0].a;
    for (const Span &sp : spans) {
      Iv v;
      v.cls = sp.cls;
      if (cudaEventElapsedTime(&v.a, ref, sp.a) != cudaSuccess || cudaEventElapsedTime(&v.b, ref, sp.b) != cudaSuccess ||
          sp.cls < 0 || sp.cls >= kKClasses) {
        cudaGetLastError();
        continue;
      }
      iv.push_back(v);
      launches[sp.cls]++;
    }
    std::sort(iv.begin(), iv.end(), [](const Iv &x, const Iv &y) { return x.a < y.a; });
    static const bool dump = env_int("MB200_TIMING_DUMP", 0) != 0; // diagnostics: every span, ms since the first event
    if (dump) {
      static const char *names[kKClasses] = {"camera_trace", "shadow_trace", "bounce_trace", "shade", "resolve", "query_trace"};
      for (const Iv &v : iv) fprintf(stderr, "[mb200 span] %-13s %9.3f -> %9.3f ms\n", names[v.cls], v.a, v.b);
    }
    auto union_len = [&](auto pred) {
      double total = 0.0;
      float lo = 0.f, hi = -1.f;
      bool open = false;
      for (const Iv &v : iv) {
        if (!pred(v.cls)) continue;
        if (!open || v.a > hi) {
          if (open) total += (double)(hi - lo);
          lo = v.a, hi = v.b, open = true;
        } else if (v.b > hi) {
          hi = v.b;
        }
      }
      if (open) total += (double)(hi - lo);
      return total;
    };
    for (int c = 0; c < kKClasses; c++) ms[c] += union_len([c](int k) { return k == c; });
    if (trace_union_ms)
      *trace_union_ms = union_len([](int k) { return k == kKCameraTrace || k == kKShadowTrace || k == kKBounceTrace; });
  }
  for (const Span &sp : spans) {
    pool.push_back(sp.a);
    pool.push_back(sp.b);
  }
  spans.clear();
}

cudaError_t FramePipe::init() {
  if (aux) return cudaSuccess;
  cudaError_t e = cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking);
  if (e != cudaSuccess) return e;
  if ((e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
  for (cudaEvent_t *ev : {&fork, &join, &resolved[0], &resolved[1]}) {
    e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  for (cudaEvent_t &ev : chunk_ev)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t FramePipe::reserve_tiles(size_t tiles, cudaStream_t s) {
  if (tiles <= tiles_cap) return cudaSuccess;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  if (hot) cudaFree(hot);
  if (order) cudaFree(order);
  hot = nullptr, order = nullptr, tiles_cap = 0, have_hot = false;
  const size_t cap = tiles + tiles / 4 + 1024;
  if ((e = cudaMalloc(&hot, cap)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&order, cap * sizeof(uint32_t))) != cudaSuccess) return e;
  tiles_cap = cap;
  return cudaSuccess;
}

void FramePipe::release() {
  for (cudaEvent_t ev : {fork, join, resolved[0], resolved[1]})
    if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t &ev : chunk_ev) {
    if (ev) cudaEventDestroy(ev);
    ev = nullptr;
  }
  if (copy) cudaStreamDestroy(copy);
  copy = nullptr;
  if (aux) cudaStreamDestroy(aux);
  if (hot) cudaFree(hot);
  if (order) cudaFree(order);
  aux = nullptr;
  fork = join = resolved[0] = resolved[1] = nullptr;
  hot = nullptr, order = nullptr, tiles_cap = 0, have_hot = false;
}

void KernelTimer::release() {
  for (const Span &sp : spans) cudaEventDestroy(sp.a), cudaEventDestroy(sp.b);
  for (cudaEvent_t e : pool) cudaEventDestroy(e);
  spans.clear();
  pool.clear();
}

namespace {
struct TimedScope { // brackets one kernel launch
  KernelTimer *t;
  cudaStream_t s;
  TimedScope(KernelTimer *timer, int cls, cudaStream_t stream) : t(timer), s(stream) {
    if (t) t->begin(cls, s);
  }
  ~TimedScope() {
    if (t) t->end(s);
  }
};
} // namespace

int band_rows_owned(int rows, int band_rows, int count, int index) {
  return band_local_rows(rows, band_rows, count, index);
}

cudaError_t launch_trace_closest(const SceneView &sc, int stack_cap, const mb200_ray *rays, size_t n,
                                 mb200_hit *hits, unsigned long long *work, unsigned long long *counters,
                                 cudaStream_t s, KernelTimer *timer) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  const IOClosest io{{0u}, rays, hits};
  TimedScope ts(timer, kKQueryTrace, s);
  return launch_trace<IOClosest, false>(sc, stack_cap, io, n, nullptr, work, counters, s);
}

cudaError_t launch_trace_occluded(const SceneView &sc, int stack_cap, const mb200_ray *rays, const double *tmax,
                                  size_t n, unsigned char *occ, unsigned long long *work,
                                  unsigned long long *counters, cudaStream_t s, KernelTimer *timer) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  const IOOccluded io{{0u}, rays, tmax, occ};
  TimedScope ts(timer, kKQueryTrace, s);
  return launch_trace<IOOccluded, true>(sc, stack_cap, io, n, nullptr, work, counters, s);
}

cudaError_t launch_build_isects(const SceneView &sc, const mb200_ray *rays, const mb200_hit *hits, size_t n,
                                mb200_isect *isects, unsigned char *mask, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_build_isects<<<flat_grid(n, 256), 256, 0, s>>>(sc, rays, hits, n, isects, mask);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_generate_rays(const mb200_camera_frame &f, const double *px, const double *py, size_t n,
                                 mb200_ray *rays, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_generate_rays<<<flat_grid(n, 256), 256, 0, s>>>(f, px, py, n, rays);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_generate_rays_env(const double origin[3], int width, int height, int stereo, const double *px,
                                     const double *py, size_t n, mb200_ray *rays, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  mb200_camera_frame f;
  memset(&f, 0, sizeof(f));
  for (int c = 0; c < 3; c++) f.origin[c] = origin[c];
  k_generate_rays_env<<<flat_grid(n, 256), 256, 0, s>>>(f, width, height, stereo, px, py, n, rays);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_generate_grid(const mb200_camera_frame &f, int x0, int y0, int w, int h, mb200_ray *rays,
                                 cudaStream_t s) {
  const size_t n = (size_t)w * h;
  if (n == 0) return cudaSuccess;
  k_generate_grid<<<flat_grid(n, 256), 256, 0, s>>>(f, x0, y0, w, h, rays);
  g_launches++;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// frame scratch + the wavefront driver
// ---------------------------------------------------------------------------
cudaError_t frame_scratch_reserve(FrameScratch &fs, size_t bytes, cudaStream_t s) {
  if (fs.bytes >= bytes) return cudaSuccess;
  if (fs.base) {
    cudaError_t e = cudaStreamSynchronize(s); // earlier work may still use the old block
    if (e != cudaSuccess) return e;
    cudaFree(fs.base);
    fs.base = nullptr, fs.bytes = 0;
  }
  const size_t cap = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&fs.base, cap);
  if (e != cudaSuccess) {
    fs.base = nullptr;
    return e;
  }
  fs.bytes = cap;
  return cudaSuccess;
}

void frame_scratch_release(FrameScratch &fs) {
  if (fs.base) cudaFree(fs.base);
  fs.base = nullptr, fs.bytes = 0;
}

// Work items per batch: bounds the scratch (100 B/item for primary+shadow, 212 B/item for PathTrace).
static size_t batch_item_budget() {
  static const size_t v = (size_t)env_int("MB200_FRAME_BATCH_ITEMS", 1 << 24);
  return v < 32 ? 32 : v;
}

cudaError_t launch_frame(const SceneView &sc, int stack_cap, const mb200_render_params &p, int num_passes, int mode,
                         float *image, int *count, FrameScratch &scratch, unsigned long long *stats, cudaStream_t s,
                         KernelTimer *timer, FramePipe *pipe, FrameChunks *chunks) {
  if (chunks) chunks->n = 0;
  const FrameMap m0 = make_frame_map(p, p.pass, 1);
  const size_t tiles = frame_map_tiles(m0);
  if (tiles == 0 || num_passes < 1) return cudaSuccess;
  const bool path = (p.shader == MB200_SHADER_PATHTRACE || p.shader == MB200_SHADER_PATHTRACE_ENV) && p.max_path_length > 1;
  // primary+shadow and primary-only samples need no path state, so the lane that traces the camera ray can shade it
  // and trace its shadow ray itself (trace_sm.cuh: IOFrameFusedT; MB200_FRAME_FUSED=1).  Measured: one
  // launch per batch and no hit-record / queue traffic, but the shade step runs inside the traversal kernel with
  // ~8 of 32 lanes and costs more issue slots than the wavefront form's full-width shade kernel, which hides behind
  // the other stream's traversal launch: 23.3 vs 21.8 ms per bench frame.  The wavefront form is the default.
  static const bool fused_on = env_int("MB200_FRAME_FUSED", 0) != 0;
  const bool fused = fused_on && (p.shader == MB200_SHADER_PRIMARY_SHADOW || p.shader == MB200_SHADER_PRIMARY_ONLY);
  const bool shadow = p.shader == MB200_SHADER_PRIMARY_SHADOW && !fused;
  const size_t item_limit = fused ? 0x7FFFFFE0ull : 0xFFFFFFE0ull; // fused: bit 31 of the item marks the shadow ray
  if (tiles * 32 > item_limit) return cudaErrorInvalidValue;

  // ---- batches.  By tile rows when all passes of one tile row fit a batch (then a batch's pixels are final when it is
  // resolved and the caller can start copying them: `chunks`); by passes over the whole tile set otherwise.  Two streams
  // need at least two batches to overlap one batch's drain with the other's kernels.
  static const bool pipeline_off = env_int("MB200_FRAME_PIPELINE", 1) == 0;
  static const bool rows_off = env_int("MB200_FRAME_ROWSPLIT", 1) == 0;
  // (a single pass is cut in two only when it is big enough for the second launch to pay: 2^20 samples)
  const bool piped = pipe && !pipeline_off && (num_passes >= 2 || tiles * 32 >= ((size_t)1 << 20));
  const size_t tiles_x = (size_t)m0.tiles_x, tiles_y = tiles / tiles_x;
  const size_t row_items = tiles_x * 32 * (size_t)num_passes; // one tile row, all passes
  const size_t budget = batch_item_budget() < item_limit ? batch_item_budget() : item_limit;
  const bool by_rows = !rows_off && row_items <= budget && tiles_y >= 2 && piped;
  struct Batch {
    uint32_t tile0, ntiles;
    int pass0, npasses;
  };
  std::vector<Batch> batches;
  size_t max_items = 0;
  int rows_per_batch = 0;
  if (by_rows) {
    size_t rows_max = budget / row_items; // tile rows per batch
    size_t nb = (tiles_y + rows_max - 1) / rows_max;
    const size_t want = (chunks && chunks->want > 2) ? (size_t)chunks->want : 2;
    if (nb < want) nb = want < tiles_y ? want : tiles_y;
    rows_per_batch = (int)((tiles_y + nb - 1) / nb);
    for (size_t r = 0; r < tiles_y; r += (size_t)rows_per_batch) {
      const size_t rows = (r + rows_per_batch <= tiles_y) ? (size_t)rows_per_batch : tiles_y - r;
      batches.push_back({(uint32_t)(r * tiles_x), (uint32_t)(rows * tiles_x), 0, num_passes});
      if (rows * row_items > max_items) max_items = rows * row_items;
    }
  } else {
    size_t per_batch = budget / (tiles * 32);
    if (per_batch < 1) per_batch = 1;
    if (per_batch > (size_t)num_passes) per_batch = (size_t)num_passes;
    if (piped && num_passes >= 2 && per_batch > (size_t)(num_passes + 1) / 2) per_batch = (size_t)(num_passes + 1) / 2;
    for (int done = 0; done < num_passes; done += (int)per_batch) {
      const int nb = (int)((size_t)(num_passes - done) < per_batch ? (size_t)(num_passes - done) : per_batch);
      batches.push_back({0u, (uint32_t)tiles, done, nb});
    }
    max_items = tiles * 32 * per_batch;
  }
  const bool two_streams = piped && batches.size() >= 2;

  // scratch carve-up (one slot per stream)
  const size_t n_traces = path ? (size_t)p.max_path_length : 2;
  const size_t ctl_bytes = align_up(sizeof(unsigned long long) * (16 + n_traces) + sizeof(unsigned int) * (n_traces + 1), 256);
  const size_t hits_bytes = fused ? 0 : align_up(max_items * sizeof(mb200_hit), 256);
  const size_t contrib_bytes = align_up(max_items * sizeof(float), 256);
  const size_t queue_bytes = (path || shadow) ? align_up(max_items * sizeof(QRay), 256) : 0;
  const size_t state_bytes = path ? align_up(max_items * sizeof(PathState), 256) : 0;
  // bounce queues are regrouped by octant and origin cell before they are traced (k_sort_*), MB200_SORT_BOUNCES=0: not
  // 2 (default): grouped by octant inside every CTA of the shade kernels, no extra pass; 1: full counting sort; 3: both
  static const int sort_mode = env_int("MB200_SORT_BOUNCES", 2);
  const bool sort_q = path && (sort_mode & 1), group_q = path && (sort_mode & 2);
  const size_t sort_bytes = sort_q ? align_up(kSortBuckets * sizeof(unsigned int), 256) : 0;
  const size_t slot_bytes = ctl_bytes + hits_bytes + contrib_bytes + queue_bytes * (path ? 2 : 1) + state_bytes + sort_bytes;
  SortGrid sgrid;
  for (int k = 0; k < 3; k++) {
    const double ext = sc.root_box[3 + k] - sc.root_box[k];
    sgrid.lo[k] = sc.root_box[k];
    sgrid.scale[k] = (ext > 0.0 && ext < DBL_MAX) ? 16.0 / ext : 0.0;
  }
  cudaError_t e = cudaSuccess;
  if (two_streams || (pipe && chunks)) {
    if ((e = pipe->init()) != cudaSuccess) return e;
    // the previous frame may still be running on aux when the block has to grow
    if (scratch.bytes < slot_bytes * 2 && (e = cudaStreamSynchronize(pipe->aux)) != cudaSuccess) return e;
  }
  if ((e = frame_scratch_reserve(scratch, slot_bytes * (two_streams ? 2 : 1), s)) != cudaSuccess) return e;
  // longest-rays-first schedule from the previous frame's flags (same tile layout and batch cut only); the order is a
  // permutation inside every batch's tile range
  const uint32_t *order = nullptr;
  unsigned char *hot = nullptr;
  static const bool lpt_off = env_int("MB200_FRAME_LPT", 1) == 0;
  if (pipe && !lpt_off && tiles >= 1024) {
    const long long layout[12] = {p.width, p.height, p.x0, p.y0, p.x1, p.y1, p.band_rows * 1000003LL + p.band_count,
                                  p.band_index, p.band_compact, p.pixel_step, (long long)tiles,
                                  by_rows ? rows_per_batch : 0};
    if ((e = pipe->reserve_tiles(tiles, s)) != cudaSuccess) return e;
    const bool same = pipe->have_hot && memcmp(layout, pipe->layout, sizeof(layout)) == 0;
    if (same) {
      // one launch, one CTA per tile range (the batches of a row cut; the whole tile set otherwise)
      for (size_t b0 = 0; b0 < (by_rows ? batches.size() : 1); b0 += 64) {
        OrderRanges rg;
        unsigned nr = 0;
        if (by_rows)
          for (size_t b = b0; b < batches.size() && nr < 64; b++, nr++) rg.tile0[nr] = batches[b].tile0, rg.tiles[nr] = batches[b].ntiles;
        else
          rg.tile0[0] = 0u, rg.tiles[0] = (uint32_t)tiles, nr = 1;
        k_build_order<<<nr, 1024, 0, s>>>(pipe->hot, rg, pipe->order);
        g_launches++;
      }
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      order = pipe->order;
    } else {
      if ((e = cudaMemsetAsync(pipe->hot, 0, tiles, s)) != cudaSuccess) return e;
      memcpy(pipe->layout, layout, sizeof(layout));
    }
    hot = pipe->hot;
    pipe->have_hot = true;
  }
  if (two_streams) { // aux starts after whatever the caller queued on s (uploads, the previous frame)
    if ((e = cudaEventRecord(pipe->fork, s)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(pipe->aux, pipe->fork, 0)) != cudaSuccess) return e;
  }

  // which batches end a chunk the caller is told about (row cut, compact or unbanded buffers only)
  const bool report = chunks && by_rows && pipe && (p.band_rows == 0 || p.band_compact);
  const size_t nbatch = batches.size();
  const size_t nchunks = report ? (nbatch < (size_t)kMaxFrameChunks ? nbatch : (size_t)kMaxFrameChunks) : 0;
  const size_t per_chunk = nchunks ? (nbatch + nchunks - 1) / nchunks : 0;
  const int step = m0.step;

  bool aux_used = false;
  for (size_t bi = 0; bi < nbatch; bi++) {
    const Batch &B = batches[bi];
    const int slot = two_streams ? (int)(bi & 1) : 0;
    const cudaStream_t st = slot ? pipe->aux : s;
    aux_used |= slot != 0;
    char *base = reinterpret_cast<char *>(scratch.base) + (size_t)slot * slot_bytes;
    unsigned long long *bstats = reinterpret_cast<unsigned long long *>(base);           // [8]
    unsigned long long *tcount = bstats + 8;  // [8] traversal counters: fused launch, or camera [0..3] + shadow [4..7]
    unsigned long long *work = tcount + 8;                                               // [n_traces]
    unsigned int *qcount = reinterpret_cast<unsigned int *>(work + n_traces);            // [n_traces + 1]
    mb200_hit *hits = reinterpret_cast<mb200_hit *>(base + ctl_bytes);
    float *contrib = reinterpret_cast<float *>(base + ctl_bytes + hits_bytes);
    QRay *queue[2] = {reinterpret_cast<QRay *>(base + ctl_bytes + hits_bytes + contrib_bytes), nullptr};
    queue[1] = path ? reinterpret_cast<QRay *>(reinterpret_cast<char *>(queue[0]) + queue_bytes) : queue[0];
    PathState *states = path ? reinterpret_cast<PathState *>(reinterpret_cast<char *>(queue[0]) + 2 * queue_bytes) : nullptr;
    unsigned int *sort_hist = sort_q ? reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(queue[0]) + 2 * queue_bytes + state_bytes) : nullptr;

    FrameMap m = make_frame_map(p, p.pass + (uint32_t)B.pass0, (uint32_t)B.npasses);
    m.order = order, m.hot = hot, m.tile0 = B.tile0;
    {
      static const int hs = env_int("MB200_HOT_STEPS", (int)kHotSteps);
      m.hot_steps = (uint32_t)hs;
    }
    const uint32_t items = (uint32_t)((size_t)B.ntiles * 32 * (size_t)B.npasses);
    if ((e = cudaMemsetAsync(base, 0, ctl_bytes, st)) != cudaSuccess) return e;

    if (fused) {
      // one launch per batch: camera ray -> shade -> shadow ray in the lane; only contrib[] leaves the kernel
      TimedScope ts(timer, kKCameraTrace, st);
      if (p.camera_mode == MB200_CAMERA_PINHOLE) {
        const IOFrameFusedT<true> io{p, m, contrib};
        e = stats ? launch_trace_kind<IOFrameFusedT<true>, false, true>(sc, stack_cap, io, items, nullptr, work + 0, tcount, st)
                  : launch_trace_kind<IOFrameFusedT<true>, false, false>(sc, stack_cap, io, items, nullptr, work + 0, nullptr, st);
      } else {
        const IOFrameFusedT<false> io{p, m, contrib};
        e = stats ? launch_trace_kind<IOFrameFusedT<false>, false, true>(sc, stack_cap, io, items, nullptr, work + 0, tcount, st)
                  : launch_trace_kind<IOFrameFusedT<false>, false, false>(sc, stack_cap, io, items, nullptr, work + 0, nullptr, st);
      }
      if (e != cudaSuccess) return e;
    } else {
      // camera rays: K1 fused into K2
      {
        TimedScope ts(timer, kKCameraTrace, st);
        // a frame whose caller asked for stats also counts box / triangle tests (primary+shadow frames only)
        unsigned long long *cc = (stats && !path) ? tcount : nullptr;
        if (p.camera_mode == MB200_CAMERA_PINHOLE) {
          const IOCameraT<true> cam{p, m, hits};
          e = launch_trace<IOCameraT<true>, false>(sc, stack_cap, cam, items, nullptr, work + 0, cc, st);
        } else {
          const IOCameraT<false> cam{p, m, hits};
          e = launch_trace<IOCameraT<false>, false>(sc, stack_cap, cam, items, nullptr, work + 0, cc, st);
        }
      }
      if (e != cudaSuccess) return e;
      {
        TimedScope ts(timer, kKShade, st);
        k_shade_primary<<<(items + 255) / 256, 256, 0, st>>>(sc, p, m, items, hits, contrib, queue[0], qcount + 0, states, bstats,
                                                             group_q ? 1 : 0);
      }
      g_launches++;
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }

    if (shadow) {
      const IOQueueShadow io{queue[0], contrib, m};
      TimedScope ts(timer, kKShadowTrace, st);
      if ((e = launch_trace<IOQueueShadow, true>(sc, stack_cap, io, 0, qcount + 0, work + 1, stats ? tcount + 4 : nullptr, st)) != cudaSuccess) return e;
    } else if (path) {
      for (int len = 2; len <= p.max_path_length; len++) {
        int qi = len & 1; // segment `len` reads queue[qi], writes queue[qi ^ 1]
        if (sort_q) { // the shade kernels always fill queue[0]; its regrouped copy in queue[1] is what is traced and shaded
          TimedScope ts(timer, kKShade, st);
          if ((e = cudaMemsetAsync(sort_hist, 0, kSortBuckets * sizeof(unsigned int), st)) != cudaSuccess) return e;
          k_sort_count<<<num_sms() * 4, 256, 0, st>>>(sgrid, queue[0], qcount + (len - 2), sort_hist);
          k_sort_scan<<<1, 1024, 0, st>>>(sort_hist);
          k_sort_scatter<<<num_sms() * 4, 256, 0, st>>>(sgrid, queue[0], qcount + (len - 2), sort_hist, queue[1]);
          g_launches += 3;
          if ((e = cudaGetLastError()) != cudaSuccess) return e;
          qi = 1;
        }
        const IOQueueClosest io{queue[qi], hits, m};
        {
          TimedScope ts(timer, kKBounceTrace, st);
          e = launch_trace_nocount<IOQueueClosest, false>(sc, stack_cap, io, 0, qcount + (len - 2), work + (len - 1), st);
        }
        if (e != cudaSuccess) return e;
        TimedScope ts(timer, kKShade, st);
        k_shade_bounce<<<num_sms() * 8, 256, 0, st>>>(sc, p, (unsigned int)len, queue[qi], qcount + (len - 2), hits, queue[qi ^ 1],
                                                      qcount + (len - 1), states, contrib, bstats, group_q ? 1 : 0);
        g_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
      }
    }

    // pass cut: the image is accumulated in pass order (float += float), batch k resolves after batch k - 1.
    // row cut: every batch holds all passes of its own pixels, nothing to order.
    int bmode = mode;
    if (!by_rows) {
      if (two_streams && bi > 0 && (e = cudaStreamWaitEvent(st, pipe->resolved[(bi - 1) & 1], 0)) != cudaSuccess) return e;
      if (mode == 2 && B.pass0 > 0) bmode = 1;
    }
    {
      TimedScope ts(timer, kKResolve, st);
      k_resolve<<<(unsigned)(((size_t)B.ntiles * 32 + 255) / 256), 256, 0, st>>>(m, B.ntiles, bmode, contrib, image, count);
    }
    g_launches++;
    if (!by_rows && two_streams && (e = cudaEventRecord(pipe->resolved[bi & 1], st)) != cudaSuccess) return e;
    if (nchunks) { // the last batch of a chunk on each stream carries one of the chunk's two events
      const size_t c = bi / per_chunk, last = ((c + 1) * per_chunk < nbatch ? (c + 1) * per_chunk : nbatch) - 1;
      if (bi == last || (bi + 1 == last && two_streams)) {
        cudaEvent_t ev = pipe->chunk_ev[2 * c + (bi == last ? 0 : 1)];
        if ((e = cudaEventRecord(ev, st)) != cudaSuccess) return e;
        if (bi == last) {
          const size_t first = c * per_chunk;
          const int r0 = (int)(batches[first].tile0 / tiles_x) * 4 * step;
          int r1 = (int)((B.tile0 + B.ntiles) / tiles_x) * 4 * step;
          const int rows_buf = (p.band_rows > 0) ? m0.rows_local : (p.y1 - p.y0);
          if (r1 > rows_buf) r1 = rows_buf;
          const int off = (p.band_rows > 0) ? 0 : p.y0; // unbanded buffers are indexed by image row
          chunks->row0[c] = off + r0, chunks->row1[c] = off + r1;
          chunks->done_a[c] = ev;
          if (!(two_streams && last > first)) chunks->done_b[c] = nullptr;
          chunks->n = (int)c + 1;
        } else {
          chunks->done_b[c] = ev;
        }
      }
    }
    if (stats) {
      if (fused) k_add_stats_fused<<<1, 32, 0, st>>>(tcount, stats);
      else k_add_stats<<<1, 32, 0, st>>>(bstats, shadow ? qcount : nullptr, tcount, stats);
      g_launches++;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (two_streams && aux_used) { // everything the frame queued on aux is ordered before what follows on s
    if ((e = cudaEventRecord(pipe->join, pipe->aux)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(s, pipe->join, 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_resolve_ldr(const float *image, const int *count, size_t npix, int mode, unsigned char *out, cudaStream_t s) {
  if (npix == 0) return cudaSuccess;
  k_resolve_ldr<<<flat_grid(npix, 256), 256, 0, s>>>(image, count, npix, mode, out);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_octant_nodes(const PairNode *src, size_t n, PairNode *dst, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_octant_nodes<<<(unsigned)((n * 8 + 255) / 256), 256, 0, s>>>(src, (uint32_t)n, dst);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_pack_nodes64(const PairNode *src, size_t n, PairNode64 *dst, int *bad, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_pack_nodes64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, (uint32_t)n, dst, bad);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_pad_tris(const void *src, int src_f32, size_t n, int kind, void *dst, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_pad_tris<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, src_f32, (uint32_t)n, kind, dst);
  g_launches++;
  return cudaGetLastError();
}

} // namespace mb200

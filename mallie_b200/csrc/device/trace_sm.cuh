// trace_sm.cuh -- the persistent-warp traversal state machine (K2 closest hit / K4 any hit), and its
// fused form that carries a sample from its camera ray through shading to its shadow ray in one lane.
//
// Why a state machine.  BVHAccel::Traverse (bvh_accel.cc:773-844) alternates two very different
// pieces of arithmetic per ray -- slab tests on inner nodes and Moeller-Trumbore on leaf triangles --
// in a data-dependent order, and rays finish after very different amounts of work.  Run as a plain
// per-thread loop over a fixed batch of 32 rays, a warp drains down to its slowest ray and runs the
// whole 1..15-triangle loop of a leaf for whichever lanes happen to be there: measured on the
// 1 M-triangle scene that kernel executed the triangle code with 4 of 32 lanes active and the whole
// kernel with 8 (profiles/r1_trace_baseline.md).  Here every lane owns one ray and is in one of
// these states,
//     INNER  (rc == kBranch)   next step = one PairNode visit (two slab tests)
//     LEAF   (1 <= rc < kShade) next step = ONE triangle test of the rc left in the leaf
//     SHADE  (rc == kShade)    fused frames only: the lane's camera ray is finished and waits for the
//                              warp's next shade step (BuildIntersection + the shadow-ray set-up)
//     IDLE   (rc == kIdle)     no ray
// and every warp iteration performs at most one step per lane (both step bodies every iteration, each
// under its own predicate), so no lane ever waits for another lane's leaf loop.  Lanes whose ray has
// finished are refilled from a per-warp pool of consecutive ray indices (one atomicAdd per CHUNK
// rays), compacted with __ballot_sync/__popc over the idle lanes, so a warp never drains.
//
// Fused frames (IO::kFused; render.cc:401-426 for one sample).  The wavefront form of a primary+shadow
// frame wrote a 32-byte hit record per camera ray, re-read it in a shade kernel that regenerated the
// camera ray, queued a 64-byte shadow ray and traced the queue in a second launch.  Here the lane
// that traced the camera ray keeps it: accepted hits go to a 32-byte per-lane slot in shared memory,
// a finished camera ray parks in SHADE, and once enough lanes of the warp are parked (or idle) one
// shade step runs BuildIntersection and the shadow-ray set-up for all of them; the lane then walks
// the tree again as an any-hit ray and deposits its sample's contribution.  Nothing but contrib[item]
// (4 bytes) leaves the kernel.
//
// Exactness.  The per-ray sequence of node visits, triangle tests and pop-time culling decisions is
// the reference's (proof sketch in traverse.cuh); only the interleaving BETWEEN rays changes, and
// rays do not interact.  Hit records stay bit-identical.
//
// Any hit, exactly.  "occluded" is DEFINED as "closest-hit Traverse returns t < tmax".  An any-hit ray
// therefore runs the closest-hit walk unchanged (hitT starts at DBL_MAX, every accepted triangle lowers
// it, boxes are culled against it) and stops at the first accepted t < tmax: up to that point its state
// equals the closest-hit walk's, whose hitT only decreases afterwards, so the two agree on every input.
// (Round 1 started hitT at tmax, which also culls boxes beyond tmax; that is the same predicate only
// while the boxes' kEPS padding dominates the slab test's rounding, i.e. not for coordinates >~ 1e3.)
#ifndef MALLIE_B200_TRACE_SM_CUH_
#define MALLIE_B200_TRACE_SM_CUH_

#include "shade.cuh"
#include "traverse.cuh"

namespace mb200 {

constexpr uint32_t kIdle = 0xFFFFFFFEu;  // lane holds no ray
constexpr uint32_t kShade = 0xFFFFFFFDu; // fused frames: camera ray finished, waiting for the shade step
constexpr uint32_t kFullMask = 0xFFFFFFFFu;
constexpr uint32_t kShadowBit = 0x80000000u; // fused frames: bit 31 of `item` = the lane's ray is the sample's shadow ray

__device__ __forceinline__ void store_hit(mb200_hit *dst, double t, double u, double v, uint32_t face, uint32_t mat) {
  double2 *o = reinterpret_cast<double2 *>(dst);
  o[0] = make_double2(t, u);
  o[1] = make_double2(v, __longlong_as_double((long long)(((unsigned long long)mat << 32) | face)));
}
// Miss record: t = DBL_MAX, u = v = 0, faceID = materialID = ~0 (bvh_accel.cc:783-786)
__device__ __forceinline__ void store_miss(mb200_hit *dst) { store_hit(dst, DBL_MAX, 0.0, 0.0, 0xFFFFFFFFu, 0xFFFFFFFFu); }

// ---- ray sources / result sinks -------------------------------------------------------------------
// load(i, ...) returns false when item i carries no ray; otherwise it yields the ray and tmax (any-hit
// sources; ignored for closest hit) and writes the "miss" result.  accept() is called on every accepted
// triangle (a handful per ray: traversal is front to back) and overwrites the result in place, so u, v,
// faceID and materialID never occupy registers between steps.  finish() ends the ray.
struct NoCostMap {
  uint32_t hot_steps;
};

// K2 over a caller's ray buffer (mb200_trace_closest).
struct IOClosest {
  static constexpr bool kTracksCost = false, kFused = false;
  NoCostMap m;
  __device__ __forceinline__ void mark_hot(uint32_t) const {}
  const mb200_ray *rays;
  mb200_hit *hits;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &tmax) const {
    const double2 *p = reinterpret_cast<const double2 *>(rays + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    tmax = DBL_MAX;
    store_miss(hits + i);
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t i, double t, double u, double v, uint32_t face, uint32_t mat) const {
    store_hit(hits + i, t, u, v, face, mat);
  }
  __device__ __forceinline__ void finish(uint32_t, bool) const {}
};

// K4 over a caller's ray buffer (mb200_trace_occluded).
struct IOOccluded {
  static constexpr bool kTracksCost = false, kFused = false;
  NoCostMap m;
  __device__ __forceinline__ void mark_hot(uint32_t) const {}
  const mb200_ray *rays;
  const double *tmax;
  unsigned char *occluded;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &t0) const {
    const double2 *p = reinterpret_cast<const double2 *>(rays + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    t0 = __ldg(tmax + i);
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t, double, double, double, uint32_t, uint32_t) const {}
  __device__ __forceinline__ void finish(uint32_t i, bool occ) const { occluded[i] = occ ? 1 : 0; }
  // re-read on the (rare) accepted triangle instead of occupying two registers for the whole walk
  __device__ __forceinline__ double tmax_of(uint32_t i) const { return __ldg(tmax + i); }
};

// K1 fused into K2: the camera ray of work item i (one jittered sample of one pixel) is generated in
// the refill step -- Camera::GenerateRay (camera.cc:222-240) after PathTrace's jitter (render.cc:386-393)
// -- and never stored; hits[i] receives the 32-byte record.  (The wavefront frames: PathTrace.)
// PINHOLE: the kernel is compiled for Camera::GenerateRay only.  The panorama cameras (sin / cos / fmod / atan2 in
// double: ~1 400 SASS instructions) get their own instantiation, which keeps the pinhole kernels inside the
// 32 KB instruction cache level (kernels beyond ~2 048 instructions lose 10-30 %).
template <bool PINHOLE> struct IOCameraT {
  static constexpr bool kTracksCost = true, kFused = false;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, i); }
  mb200_render_params p;
  FrameMap m;
  mb200_hit *hits;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &tmax) const {
    int x, y, rl;
    uint32_t pass;
    store_miss(hits + i);
    if (!item_pixel(m, i, x, y, rl, pass)) return false;
    Xorshift128 rng;
    camera_sample<PINHOLE>(p, x, y, pass, rng, ox, oy, oz, dx, dy, dz);
    tmax = DBL_MAX;
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t i, double t, double u, double v, uint32_t face, uint32_t mat) const {
    store_hit(hits + i, t, u, v, face, mat);
  }
  __device__ __forceinline__ void finish(uint32_t, bool) const {}
};

// K2 over a queue of path-continuation rays: hits[i] for queue slot i.
struct IOQueueClosest {
  static constexpr bool kTracksCost = true, kFused = false;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, __ldg(&q[i].item)); }
  const QRay *q;
  mb200_hit *hits;
  FrameMap m;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &tmax) const {
    const double2 *p = reinterpret_cast<const double2 *>(q + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    tmax = DBL_MAX;
    store_miss(hits + i);
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t i, double t, double u, double v, uint32_t face, uint32_t mat) const {
    store_hit(hits + i, t, u, v, face, mat);
  }
  __device__ __forceinline__ void finish(uint32_t, bool) const {}
};

// K4 over the queue of shadow rays: an unoccluded ray deposits its `value` into its sample's slot.
// (The wavefront form of the primary+shadow frame, kept for A/B runs: MB200_FRAME_FUSED=0.)
struct IOQueueShadow {
  static constexpr bool kTracksCost = true, kFused = false;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, __ldg(&q[i].item)); }
  const QRay *q;
  float *contrib; // [items]
  FrameMap m;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &t0) const {
    const double2 *p = reinterpret_cast<const double2 *>(q + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    t0 = d.x;
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t, double, double, double, uint32_t, uint32_t) const {}
  __device__ __forceinline__ void finish(uint32_t i, bool occ) const {
    if (!occ) {
      const uint2 w = __ldg(reinterpret_cast<const uint2 *>(&q[i].item));
      contrib[w.x] = __uint_as_float(w.y);
    }
  }
  __device__ __forceinline__ double tmax_of(uint32_t i) const { return __ldg(&q[i].tmax); }
};

// The lane's 32-byte slot in shared memory (fused frames), two 16-byte units in the column layout of the
// traversal stack (unit k of thread t at [k * blockDim.x + t]: conflict-free 128-bit accesses):
//   while the camera ray is traced:  unit 0 = u, v   unit 1 = faceID, materialID   of the closest hit so far
//   while the shadow ray is traced:  unit 0 = tmax, -  unit 1 = value (float bits)
struct LaneSlot {
  uint32_t addr0, addr1; // shared-window addresses of the two units
  __device__ __forceinline__ void put_uv(double u, double v) const {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr0), "d"(u), "d"(v) : "memory");
  }
  __device__ __forceinline__ void get_uv(double &u, double &v) const {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(u), "=d"(v) : "r"(addr0) : "memory");
  }
  __device__ __forceinline__ void put_ids(uint32_t a, uint32_t b) const {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr1), "r"(a), "r"(b) : "memory");
  }
  __device__ __forceinline__ void get_ids(uint32_t &a, uint32_t &b) const {
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr1) : "memory");
  }
  __device__ __forceinline__ void put_tmax(double t) const {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr0), "d"(t) : "memory");
  }
  __device__ __forceinline__ double get_tmax() const {
    double t;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr0) : "memory");
    return t;
  }
};

// Fused primary+shadow / primary-only frame: camera ray -> shade -> shadow ray in the lane.
// shade() restates the body of k_shade_primary (kernels.cu) for the two shaders that need no path state.
template <bool PINHOLE> struct IOFrameFusedT {
  static constexpr bool kTracksCost = true, kFused = true;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, i & ~kShadowBit); }
  mb200_render_params p;
  FrameMap m;
  float *contrib; // [items]
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &tmax) const {
    int x, y, rl;
    uint32_t pass;
    if (!item_pixel(m, i, x, y, rl, pass)) {
      contrib[i] = 0.f; // padding lane of a ragged tile: the resolve kernel never reads it, keep it defined
      return false;
    }
    Xorshift128 rng;
    camera_sample<PINHOLE>(p, x, y, pass, rng, ox, oy, oz, dx, dy, dz);
    tmax = DBL_MAX;
    return true;
  }
  // The sample's camera ray (o, d) is finished with closest hit (t, u, v, face, mat) or face == ~0.
  // Returns true when a shadow ray has to be traced: (qo, qd, tmax) and the value it deposits if unoccluded;
  // otherwise `out` is the sample's contribution.
  __device__ __forceinline__ bool shade(const SceneView &sc, double ox, double oy, double oz, double dx, double dy,
                                        double dz, double t, double u, double v, uint32_t face, uint32_t mat,
                                        double &qox, double &qoy, double &qoz, double &qdx, double &qdy, double &qdz,
                                        double &qtmax, float &value, float &out) const {
    bool hit = face != 0xFFFFFFFFu;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    uint32_t cur_mat = 0; // Intersection::materialID before the first Trace (zero-initialised isect)
    out = 0.f;
    if (hit) {
      IsectD d;
      build_intersection(sc, ox, oy, oz, dx, dy, dz, t, u, v, face, d);
      nx = d.nx, ny = d.ny, nz = d.nz;
      cur_mat = mat;
    }
    if (p.use_plane) hit |= plane_intersect(p.plane, ox, oy, oz, dx, dy, dz, t, nx, ny, nz, cur_mat);
    if (!hit) return false; // a camera ray that escapes contributes nothing (render.cc:409-412)
    if (p.shader == MB200_SHADER_PRIMARY_ONLY) {
      out = 1.f;
      return false;
    }
    // the next-event-estimation block the reference leaves empty (render.cc:425-426)
    const double hx = ox + t * dx, hy = oy + t * dy, hz = oz + t * dz;
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
    return true;
  }
};

// ---- the state machine ------------------------------------------------------------------------------
// REFILL_MIN: idle (+ parked) lanes that trigger a refill / shade step; CHUNK: ray indices taken from the
// global counter per atomicAdd (32 keeps the end-of-launch imbalance small).
// VAR: bit set of code-generation variants for A/B runs (development builds pick them with MB200_TRACE_VAR;
// production value kVar in kernels.cu).  Round-1 variants that lost (software prefetch, vote policies, 2/4
// triangles per leaf step, single pop, branch-free triangle test, per-chunk tile decode, K rays per lane) are
// gone from the source; their logs are profiles/r1_ab*.log and DESIGN.md §5 lists them.
//   1  inner step: when all lanes in the step share the direction-sign mask, a copy of the two slab tests
//      specialised for that octant runs (no per-axis selects)
//   2  the top of the tree (SceneView::top_nodes, breadth-first, MB200_TOP_NODES=K) is staged into shared memory by
//      every CTA with cp.async.bulk (one 128-byte bulk copy per node into a 144-byte slot: lanes that read different
//      nodes then hit different banks) and inner steps whose ref carries kTopBit read it with LDS.128; the traversal
//      stack moves to local memory to make room
//   4  64-byte pair nodes (layout.h: PairNode64): two 256-bit loads per visit, the exact doubles rebuilt with twelve
//      conversions and twelve additions
//   8  octant copies of the pair nodes (layout.h: nodes_oct): boxes pre-ordered (near, far) per axis and children
//      pre-ordered (near, far) for the ray's octant, so the step has no sign selects and no axis lookup
//  16  phase vote: an iteration runs only the step body (INNER or LEAF) that more lanes of the warp wait for; the
//      other lanes keep their state.  +32: both bodies run when the smaller group has at least kVoteBoth lanes
constexpr int kVarOctant = 1, kVarTopSmem = 2, kVarNode64 = 4, kVarOctNodes = 8, kVarVote = 16, kVarVoteBoth = 32;
constexpr int kVoteBoth = 10;
constexpr uint32_t kTopSlotBytes = 144;

// Stages sc.top_nodes into shared memory at `table` (shared-window address; 16-byte aligned) and waits for it.
// mbar: 8 bytes of shared memory for the transaction barrier.
__device__ __forceinline__ void stage_top_nodes(const SceneView &sc, uint32_t mbar, uint32_t table) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(sc.top_count * 128u) : "memory");
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < sc.top_count; i += blockDim.x)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     table + i * kTopSlotBytes),
                 "l"(sc.top_nodes + i), "r"(128), "r"(mbar)
                 : "memory");
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(mbar), "r"(0)
                 : "memory");
}

__device__ __forceinline__ void lds128(uint32_t addr, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

struct NodeWords { // one PairNode as loaded
  double b[2][6];
  uint32_t ref0, ref1, cnt0, cnt1, axis;
};

// One 128-byte line as four 256-bit loads (LDG.E.256).  AXIS = false (octant copies: the children are pre-ordered): the
// last load is 128 bits wide, which keeps four destination registers free at the step's register peak.
template <bool AXIS = true> __device__ __forceinline__ NodeWords load_pair_node(const PairNode *n) {
  NodeWords w;
  const char *p = reinterpret_cast<const char *>(n);
  double a0, a1, a2, a3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p + 32 * k));
    (&w.b[0][0])[4 * k + 0] = a0, (&w.b[0][0])[4 * k + 1] = a1, (&w.b[0][0])[4 * k + 2] = a2, (&w.b[0][0])[4 * k + 3] = a3;
  }
  uint32_t m0, m1, m2, m3, m4 = 0, m5, m6, m7;
  if (AXIS) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3), "=r"(m4), "=r"(m5), "=r"(m6), "=r"(m7)
                 : "l"(p + 96));
    (void)m5, (void)m6, (void)m7;
  } else {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3) : "l"(p + 96));
  }
  w.ref0 = m0, w.ref1 = m1, w.cnt0 = m2, w.cnt1 = m3, w.axis = m4;
  return w;
}

// the 64-byte form: float f per coordinate, exact double = (double)f -/+ kEPS (layout.h)
__device__ __forceinline__ NodeWords load_pair_node64(const PairNode64 *n) {
  NodeWords w;
  const char *p = reinterpret_cast<const char *>(n);
  float f[12];
  uint32_t m0, m1, m2, m3;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(p));
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(*reinterpret_cast<uint32_t *>(&f[8])), "=r"(*reinterpret_cast<uint32_t *>(&f[9])),
                 "=r"(*reinterpret_cast<uint32_t *>(&f[10])), "=r"(*reinterpret_cast<uint32_t *>(&f[11])), "=r"(m0), "=r"(m1),
                 "=r"(m2), "=r"(m3)
               : "l"(p + 32));
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      w.b[c][k] = (double)f[6 * c + k] - MB200_TRI_EPS;         // bmin: vertex - kEPS (bvh_accel.cc:293-306)
      w.b[c][3 + k] = (double)f[6 * c + 3 + k] + MB200_TRI_EPS; // bmax: vertex + kEPS
    }
  w.ref0 = m0, w.ref1 = m1;
  const uint32_t c0 = m2 & 0xFFFFu, c1 = m2 >> 16;
  w.cnt0 = c0 == 0xFFFFu ? kBranch : c0, w.cnt1 = c1 == 0xFFFFu ? kBranch : c1;
  w.axis = m3;
  return w;
}

// the same node from the staged table (eight LDS.128)
__device__ __forceinline__ NodeWords load_pair_node_smem(uint32_t addr) {
  NodeWords w;
#pragma unroll
  for (int k = 0; k < 6; k++) lds128(addr + 16 * k, (&w.b[0][0])[2 * k], (&w.b[0][0])[2 * k + 1]);
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.ref0), "=r"(w.ref1), "=r"(w.cnt0), "=r"(w.cnt1) : "r"(addr + 96));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w.axis) : "r"(addr + 112));
  return w;
}

// counters of a launch: [0..3] nodes, tris, rays, max stack of closest-hit (camera) rays; fused frames add
// [4..6] nodes, tris, rays of the shadow rays
template <class IO, int TRI, int S, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int SHADE_MIN, unsigned CHUNK, int VAR>
__device__ __forceinline__ void trace_state_machine(const SceneView &sc, const IO &io, unsigned long long n,
                                                    unsigned long long *work, TravStack<S, CAP> &st,
                                                    const LaneSlot slot, unsigned long long *gcounters,
                                                    uint32_t top_table = 0) {
  const unsigned lane = threadIdx.x & 31u;
  // top-of-tree staging variant: the root pair is entry 0 of the table
  const uint32_t root_ref = ((VAR & kVarTopSmem) && sc.top_count && sc.root_cnt == kBranch) ? kTopBit : sc.root_ref;
  const unsigned lt_mask = (1u << lane) - 1u;

  RayD r;
  double hit_t = 0.0;
  uint32_t ref = 0, rc = kIdle, item = 0;
  int sp = 0;
  uint32_t pool_next = 0, pool_end = 0;
  bool exhausted = false;
  uint32_t iter = 0, born = 0; // warp iterations so far / at the time this lane's ray was loaded
  TravCounters cnt = {0u, 0u, 0u}, cnt_s = {0u, 0u, 0u};
  unsigned int nrays = 0, nrays_s = 0;
  r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = 0.0;
  r.sgn = 0u;

  // enters the tree at the root with the ray in r (hit_t set): INNER / LEAF state, or false when the root is missed
  auto enter_root = [&](bool shadow) -> bool {
    bool enter = false;
    if (!sc.empty) {
      double tm;
      if (COUNT) (shadow ? cnt_s.nodes : cnt.nodes)++;
      enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], r,
                        hit_t, tm);
    }
    if (enter && sc.root_cnt != 0u) {
      ref = root_ref, rc = sc.root_cnt;
      if (COUNT && rc != kBranch) (shadow ? cnt_s.tris : cnt.tris) += rc;
      return true;
    }
    return false;
  };

  for (;; iter++) {
    // ---- A. shade parked lanes, refill idle lanes ------------------------------------------------------
    // A step is worth its divergence once enough lanes take part; with nothing else left to do it runs anyway.
    const unsigned idle = __ballot_sync(kFullMask, rc == kIdle);
    unsigned parked = 0u;
    if (IO::kFused) parked = __ballot_sync(kFullMask, rc == kShade);
    if (idle | parked) {
      unsigned idle_now = idle;
      if constexpr (IO::kFused) {
        if (parked && (__popc(parked) >= SHADE_MIN || exhausted || (idle | parked) == kFullMask)) {
          if (rc == kShade) {
            double u, v, qox, qoy, qoz, qdx, qdy, qdz, qtmax;
            uint32_t face, mat;
            float value, out;
            slot.get_uv(u, v);
            slot.get_ids(face, mat);
            if (io.shade(sc, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, hit_t, u, v, face, mat, qox, qoy, qoz, qdx, qdy, qdz,
                         qtmax, value, out)) {
              ray_setup(r, qox, qoy, qoz, qdx, qdy, qdz);
              hit_t = DBL_MAX; // exact any hit: the closest-hit walk, stopped at the first t < tmax
              slot.put_tmax(qtmax);
              slot.put_ids(__float_as_uint(value), 0u);
              item |= kShadowBit;
              sp = 0;
              born = iter;
              if (COUNT) nrays_s++;
              if (!enter_root(true)) { // nothing in the way
                io.contrib[item & ~kShadowBit] = value;
                rc = kIdle;
              }
              if ((VAR & kVarOctNodes) && rc == kBranch) ref += r.sgn * sc.num_pair_nodes; // into the octant's copy
            } else {
              io.contrib[item] = out;
              rc = kIdle;
            }
          }
          idle_now = __ballot_sync(kFullMask, rc == kIdle);
          parked = 0u;
        }
      }
      if (idle_now) {
        if (!exhausted && (__popc(idle_now) >= REFILL_MIN || (idle_now | parked) == kFullMask)) {
          if (pool_next == pool_end) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work, (unsigned long long)CHUNK);
            base = __shfl_sync(kFullMask, base, 0);
            if (base >= n) {
              exhausted = true;
            } else {
              pool_next = (uint32_t)base;
              pool_end = (uint32_t)((base + CHUNK < n) ? base + CHUNK : n);
            }
          }
          if (!exhausted) {
            const unsigned avail = pool_end - pool_next, want = __popc(idle_now);
            const unsigned rank = __popc(idle_now & lt_mask);
            if (rc == kIdle && rank < avail) {
              item = pool_next + rank;
              double ox, oy, oz, dx, dy, dz, t0;
              if (io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
                ray_setup(r, ox, oy, oz, dx, dy, dz);
                hit_t = DBL_MAX; // bvh_accel.cc:774 (any hit included: see the header)
                sp = 0;
                if (IO::kTracksCost) born = iter;
                if (COUNT) nrays++;
                if (IO::kFused) slot.put_ids(0xFFFFFFFFu, 0xFFFFFFFFu); // no hit yet
                if (!enter_root(false)) {
                  if constexpr (IO::kFused) rc = kShade; // the plane may still be hit
                  else io.finish(item, false);
                }
                if ((VAR & kVarOctNodes) && rc == kBranch) ref += r.sgn * sc.num_pair_nodes; // into the octant's copy // from here on: the first node of the octant's copy
              }
            }
            pool_next += (want < avail) ? want : avail;
          }
        }
        if (exhausted && idle_now == kFullMask) break;
      }
    }

    // ---- B. INNER: one 128-byte PairNode, both children tested (equivalence: traverse.cuh) -----------------
    bool at_inner = (rc == kBranch);
    bool at_leaf = (rc - 1u) < (kShade - 1u); // 1 <= rc < kShade
    if (VAR & kVarVote) {
      const int ni = __popc(__ballot_sync(kFullMask, at_inner)), nl = __popc(__ballot_sync(kFullMask, at_leaf));
      if (!(VAR & kVarVoteBoth) || min(ni, nl) < kVoteBoth) {
        if (ni >= nl) at_leaf = false;
        else at_inner = false;
      }
    }
    if (at_inner) {
      NodeWords nw;
      if ((VAR & kVarTopSmem) && (ref & kTopBit)) nw = load_pair_node_smem(top_table + (ref & ~kTopBit) * kTopSlotBytes);
      else if (VAR & kVarNode64) nw = load_pair_node64(sc.nodes64 + ref);
      else if (VAR & kVarOctNodes) nw = load_pair_node<false>(sc.nodes_oct + ref); // branch refs of a copy are absolute
      else nw = load_pair_node(sc.nodes + ref);
      double t0, t1;
      bool h0, h1;
      bool done = false;
      if (VAR & kVarOctNodes) { // the copy is already ordered for this ray's octant
        done = true;
        h0 = slab_test_oct<0>(nw.b[0], r, hit_t, t0);
        h1 = slab_test_oct<0>(nw.b[1], r, hit_t, t1);
      }
      if (VAR & kVarOctant) {
        int uniform;
        __match_all_sync(__activemask(), r.sgn, &uniform);
        if (uniform) {
          done = true;
          h0 = h1 = false, t0 = t1 = 0.0;
#define MB200_OCT(K)                                   \
  case K:                                              \
    h0 = slab_test_oct<K>(nw.b[0], r, hit_t, t0);      \
    h1 = slab_test_oct<K>(nw.b[1], r, hit_t, t1);      \
    break;
          switch (r.sgn) { MB200_OCT(0) MB200_OCT(1) MB200_OCT(2) MB200_OCT(3) MB200_OCT(4) MB200_OCT(5) MB200_OCT(6) MB200_OCT(7) }
#undef MB200_OCT
        }
      }
      if (!done) {
        h0 = slab_test(nw.b[0][0], nw.b[0][1], nw.b[0][2], nw.b[0][3], nw.b[0][4], nw.b[0][5], r, hit_t, t0);
        h1 = slab_test(nw.b[1][0], nw.b[1][1], nw.b[1][2], nw.b[1][3], nw.b[1][4], nw.b[1][5], r, hit_t, t1);
      }
      const bool shadow = IO::kFused && (item & kShadowBit);
      if (COUNT) (shadow ? cnt_s.nodes : cnt.nodes) += 2;
      // dirSign[axis]: child 1 is the near one; the octant copies store the near child first
      const bool sgn = (VAR & kVarOctNodes) ? false : (((r.sgn >> nw.axis) & 1u) != 0u);
      if (h0 && h1) { // near = data[dirSign[axis]] first, far pushed with its tmin (bvh_accel.cc:818-823)
        st.put(sp++, sgn ? t0 : t1, sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1);
        if (COUNT) cnt.max_stack = max(cnt.max_stack, (unsigned int)sp + 1u);
        ref = sgn ? nw.ref1 : nw.ref0, rc = sgn ? nw.cnt1 : nw.cnt0;
      } else if (h0) {
        ref = nw.ref0, rc = nw.cnt0;
      } else if (h1) {
        ref = nw.ref1, rc = nw.cnt1;
      } else {
        rc = 0u;
      }
      if (COUNT && rc != 0u && rc != kBranch) (shadow ? cnt_s.tris : cnt.tris) += rc;
    } else if (at_leaf) {
      // ---- LEAF: one triangle of TestLeafNode (bvh_accel.cc:640-697), in indices_ order -----------------
      double u, v;
      bool accepted;
      uint32_t tface = 0, tmat = 0;
      if constexpr (TRI == kTriWoop) { // development variant, not bit-exact (traverse.cuh)
        const WoopRec wr = load_woop(sc.trav_tris, ref);
        accepted = tri_test_woop(hit_t, u, v, wr, r);
        if (accepted) tri_ids(sc.tris, sc.tri_f32, ref, tface, tmat);
      } else {
        const TriEdges tv = load_tri_edges<TRI>(sc.trav_tris, ref);
        accepted = tri_test_edges(hit_t, u, v, tv, r);
        tface = tv.face, tmat = tv.mat;
      }
      if (accepted) {
        bool stop = false;
        if constexpr (IO::kFused) {
          if (item & kShadowBit) {
            stop = hit_t < slot.get_tmax(); // occluded: the sample keeps contribution 0
            if (stop) io.contrib[item & ~kShadowBit] = 0.f;
          } else {
            slot.put_uv(u, v);
            slot.put_ids(tface, tmat);
          }
        } else {
          io.accept(item, hit_t, u, v, tface, tmat);
          if constexpr (ANYHIT) {
            if (hit_t < io.tmax_of(item)) { // occluded: closest-hit Traverse would return t < tmax
              io.finish(item, true);
              stop = true;
            }
          }
        }
        if (stop) {
          if (IO::kTracksCost && iter - born > io.m.hot_steps) io.mark_hot(item);
          rc = kIdle;
        }
      }
      if (rc != kIdle) {
        ref++;
        rc--;
      }
    }

    // ---- C. pop: the reference's pop-time (tmin <= hitT) decision; empty stack = ray finished ----------
    if (rc == 0u) {
      for (;;) {
        if (sp == 0) {
          if (IO::kTracksCost && iter - born > io.m.hot_steps) io.mark_hot(item);
          if constexpr (IO::kFused) {
            if (item & kShadowBit) { // unoccluded: the sample receives its value
              uint32_t vb, unused;
              slot.get_ids(vb, unused);
              io.contrib[item & ~kShadowBit] = __uint_as_float(vb);
              rc = kIdle;
            } else {
              rc = kShade;
            }
          } else {
            io.finish(item, false);
            rc = kIdle;
          }
          break;
        }
        double tm;
        st.get(--sp, tm, ref, rc);
        if (tm <= hit_t && rc != 0u) {
          if (COUNT && rc != kBranch) ((IO::kFused && (item & kShadowBit)) ? cnt_s.tris : cnt.tris) += rc;
          break;
        }
        rc = 0u;
      }
    }
  }

  if (COUNT) {
    unsigned long long a = cnt.nodes, b = cnt.tris, c = nrays, d = cnt_s.nodes, e = cnt_s.tris, f = nrays_s;
    unsigned int m = cnt.max_stack;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(kFullMask, a, o);
      b += __shfl_down_sync(kFullMask, b, o);
      c += __shfl_down_sync(kFullMask, c, o);
      m = max(m, __shfl_down_sync(kFullMask, m, o));
      if (IO::kFused) {
        d += __shfl_down_sync(kFullMask, d, o);
        e += __shfl_down_sync(kFullMask, e, o);
        f += __shfl_down_sync(kFullMask, f, o);
      }
    }
    if (lane == 0) {
      atomicAdd(&gcounters[0], a);
      atomicAdd(&gcounters[1], b);
      atomicAdd(&gcounters[2], c);
      atomicMax(&gcounters[3], (unsigned long long)m);
      if (IO::kFused) {
        atomicAdd(&gcounters[4], d);
        atomicAdd(&gcounters[5], e);
        atomicAdd(&gcounters[6], f);
      }
    }
  }
}

} // namespace mb200

#endif

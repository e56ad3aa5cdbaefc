// trace_sm.cuh -- the persistent-warp traversal state machine (K2 closest hit / K4 any hit).
//
// Why a state machine.  BVHAccel::Traverse (bvh_accel.cc:773-844) alternates two very different
// pieces of arithmetic per ray -- slab tests on inner nodes and Moeller-Trumbore on leaf triangles --
// in a data-dependent order, and rays finish after very different amounts of work.  Run as a plain
// per-thread loop over a fixed batch of 32 rays, a warp drains down to its slowest ray and runs the
// whole 1..15-triangle loop of a leaf for whichever lanes happen to be there: measured on the
// 1 M-triangle scene that kernel executed the triangle code with 4 of 32 lanes active and the whole
// kernel with 8 (profiles/r1_trace_baseline.md).  Here every lane owns one ray and is in one of
// three states,
//     INNER  (rc == kBranch)   next step = one PairNode visit (two slab tests)
//     LEAF   (1 <= rc < kIdle) next step = ONE triangle test of the rc left in the leaf
//     IDLE   (rc == kIdle)     no ray
// and every warp iteration performs at most one step per lane, so no lane ever waits for another
// lane's leaf loop.  Lanes whose ray has finished are refilled from a per-warp pool of consecutive
// ray indices (one atomicAdd per CHUNK rays), compacted with __ballot_sync/__popc over the idle
// lanes, so a warp never drains.  POLICY selects how the two step bodies are scheduled:
//     2 (production) both bodies every iteration, each under its own predicate ("if-if");
//     0              a warp vote picks the body more lanes are waiting for, the others wait
//                    (kept for A/B runs: fewer instructions, but measured slower -- DESIGN.md).
//
// Exactness.  The per-ray sequence of node visits, triangle tests and pop-time culling decisions is
// the reference's (proof sketch in traverse.cuh); only the interleaving BETWEEN rays changes, and
// rays do not interact.  Hit records stay bit-identical.
#ifndef MALLIE_B200_TRACE_SM_CUH_
#define MALLIE_B200_TRACE_SM_CUH_

#include "shade.cuh"
#include "traverse.cuh"

namespace mb200 {

constexpr uint32_t kIdle = 0xFFFFFFFEu; // lane holds no ray
constexpr uint32_t kFullMask = 0xFFFFFFFFu;

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// All 128-byte lines of a leaf's triangle records (<= 15 records of 48 or 80 bytes: at most 10 lines).
template <bool F32> __device__ __forceinline__ void prefetch_leaf(const void *tris, uint32_t ref, uint32_t cnt) {
  const size_t rec = F32 ? sizeof(TriRecordF32) : sizeof(TriRecordF64);
  const char *p = reinterpret_cast<const char *>(tris) + (size_t)ref * rec;
  const char *end = p + (size_t)cnt * rec;
  for (const char *q = reinterpret_cast<const char *>(reinterpret_cast<size_t>(p) & ~(size_t)127); q < end; q += 128)
    prefetch_l1(q);
}

__device__ __forceinline__ void store_hit(mb200_hit *dst, double t, double u, double v, uint32_t face, uint32_t mat) {
  double2 *o = reinterpret_cast<double2 *>(dst);
  o[0] = make_double2(t, u);
  o[1] = make_double2(v, __longlong_as_double((long long)(((unsigned long long)mat << 32) | face)));
}
// Miss record: t = DBL_MAX, u = v = 0, faceID = materialID = ~0 (bvh_accel.cc:783-786)
__device__ __forceinline__ void store_miss(mb200_hit *dst) { store_hit(dst, DBL_MAX, 0.0, 0.0, 0xFFFFFFFFu, 0xFFFFFFFFu); }

// ---- ray sources / result sinks -------------------------------------------------------------------
// load(i, ...) returns false when item i carries no ray; otherwise it yields the ray and the initial
// hitT and writes the "miss" result.  accept() is called on every accepted triangle (a handful per
// ray: traversal is front to back) and overwrites the result in place, so u, v, faceID and
// materialID never occupy registers between steps.  finish() ends the ray.

// Every source also has a per-thread chunk context (empty unless the source can share work between the
// rays of one 32-item chunk): prepare(ctx, base) is called by all lanes when the warp takes a new chunk.
struct NoChunk {};

// K2 over a caller's ray buffer (mb200_trace_closest).
struct NoCostMap {
  uint32_t hot_steps;
};

struct IOClosest {
  static constexpr bool kTracksCost = false;
  NoCostMap m;
  __device__ __forceinline__ void mark_hot(uint32_t) const {}
  typedef NoChunk Chunk;
  __device__ __forceinline__ void prepare(Chunk &, uint32_t) const {}
  __device__ __forceinline__ bool load(const Chunk &, uint32_t i, double &ox, double &oy, double &oz, double &dx,
                                       double &dy, double &dz, double &t0) const {
    return load(i, ox, oy, oz, dx, dy, dz, t0);
  }
  const mb200_ray *rays;
  mb200_hit *hits;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &t0) const {
    const double2 *p = reinterpret_cast<const double2 *>(rays + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    t0 = DBL_MAX; // bvh_accel.cc:783
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
  static constexpr bool kTracksCost = false;
  NoCostMap m;
  __device__ __forceinline__ void mark_hot(uint32_t) const {}
  typedef NoChunk Chunk;
  __device__ __forceinline__ void prepare(Chunk &, uint32_t) const {}
  __device__ __forceinline__ bool load(const Chunk &, uint32_t i, double &ox, double &oy, double &oz, double &dx,
                                       double &dy, double &dz, double &t0) const {
    return load(i, ox, oy, oz, dx, dy, dz, t0);
  }
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
  __device__ __forceinline__ double tmax_of(uint32_t i) const { return __ldg(tmax + i); }
};

// K1 fused into K2: the camera ray of work item i (one jittered sample of one pixel) is generated in
// the refill step -- Camera::GenerateRay (camera.cc:222-240) after PathTrace's jitter (render.cc:386-393)
// -- and never stored; hits[i] receives the 32-byte record.
struct IOCamera {
  static constexpr bool kTracksCost = true;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, i); }
  // (tile, pass) of the chunk the warp is drawing from: one decode (three integer divisions) per 32 rays
  struct Chunk {
    uint32_t group; // item >> 5 the fields below belong to (0xFFFFFFFF: none)
    int x_tile, y_tile, rows_valid;
    uint32_t pass;
  };
  __device__ __forceinline__ void prepare(Chunk &c, uint32_t base) const {
    int x, y, rl;
    item_pixel(m, base & ~31u, x, y, rl, c.pass); // lane 0 of the tile
    c.group = base >> 5;
    c.x_tile = x, c.y_tile = y;
    c.rows_valid = m.rows_local - rl; // rows of this tile that exist (>= 1)
  }
  __device__ __forceinline__ bool load(const Chunk &c, uint32_t i, double &ox, double &oy, double &oz, double &dx,
                                       double &dy, double &dz, double &t0) const {
    if ((i >> 5) != c.group) return load(i, ox, oy, oz, dx, dy, dz, t0);
    store_miss(hits + i);
    const int lx = (int)(i & 7u), ly = (int)((i >> 3) & 3u);
    const int x = c.x_tile + lx * m.step, y = c.y_tile + ly * m.step;
    if (!(x < m.x1 && ly < c.rows_valid)) return false;
    Xorshift128 rng;
    camera_sample(p, x, y, c.pass, rng, ox, oy, oz, dx, dy, dz);
    t0 = DBL_MAX;
    return true;
  }
  mb200_render_params p;
  FrameMap m;
  mb200_hit *hits;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &t0) const {
    int x, y, rl;
    uint32_t pass;
    store_miss(hits + i);
    if (!item_pixel(m, i, x, y, rl, pass)) return false;
    Xorshift128 rng;
    camera_sample(p, x, y, pass, rng, ox, oy, oz, dx, dy, dz);
    t0 = DBL_MAX;
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t i, double t, double u, double v, uint32_t face, uint32_t mat) const {
    store_hit(hits + i, t, u, v, face, mat);
  }
  __device__ __forceinline__ void finish(uint32_t, bool) const {}
};

// K2 over a queue of path-continuation rays: hits[i] for queue slot i.
struct IOQueueClosest {
  static constexpr bool kTracksCost = true;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, __ldg(&q[i].item)); }
  typedef NoChunk Chunk;
  __device__ __forceinline__ void prepare(Chunk &, uint32_t) const {}
  __device__ __forceinline__ bool load(const Chunk &, uint32_t i, double &ox, double &oy, double &oz, double &dx,
                                       double &dy, double &dz, double &t0) const {
    return load(i, ox, oy, oz, dx, dy, dz, t0);
  }
  const QRay *q;
  mb200_hit *hits;
  FrameMap m;
  __device__ __forceinline__ bool load(uint32_t i, double &ox, double &oy, double &oz, double &dx, double &dy,
                                       double &dz, double &t0) const {
    const double2 *p = reinterpret_cast<const double2 *>(q + i);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
    t0 = DBL_MAX;
    store_miss(hits + i);
    return true;
  }
  __device__ __forceinline__ void accept(uint32_t i, double t, double u, double v, uint32_t face, uint32_t mat) const {
    store_hit(hits + i, t, u, v, face, mat);
  }
  __device__ __forceinline__ void finish(uint32_t, bool) const {}
};

// K4 over the queue of shadow rays: an unoccluded ray deposits its `value` into its sample's slot.
struct IOQueueShadow {
  static constexpr bool kTracksCost = true;
  __device__ __forceinline__ void mark_hot(uint32_t i) const { mark_hot_tile(m, __ldg(&q[i].item)); }
  typedef NoChunk Chunk;
  __device__ __forceinline__ void prepare(Chunk &, uint32_t) const {}
  __device__ __forceinline__ bool load(const Chunk &, uint32_t i, double &ox, double &oy, double &oz, double &dx,
                                       double &dy, double &dz, double &t0) const {
    return load(i, ox, oy, oz, dx, dy, dz, t0);
  }
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

// ---- the state machine ------------------------------------------------------------------------------
// REFILL_MIN: idle lanes that trigger a refill; CHUNK: ray indices taken from the global counter per
// atomicAdd (32 keeps the end-of-launch imbalance small: a launch of 2 M rays is only ~18 per lane).
// VAR: bit set of code-generation variants kept for A/B runs (development builds pick them with
// MB200_TRACE_VAR; production value kVar in kernels.cu):
//   1  PairNode fetched with four 256-bit loads (LDG.E.256) instead of seven 128-bit + one 32-bit
//   2  (retired: the direction signs are always one 3-bit mask now, traverse.cuh)
//   4  leaf prefetch touches only the first 128-byte line of the leaf's records (no per-lane loop)
//   8  no software prefetch at all
//  16  branch-free triangle test (one predicate at the end instead of early returns)
//  32  at most one stack pop per warp iteration (no inner pop loop)
//  64  camera rays: the (tile, pass) decode of a 32-item chunk is done once per chunk, not per ray
// 256 / 512  a LEAF step tests up to 2 / 4 triangles of the leaf (in order) instead of one
// 1024  drain-phase prefetch: once the ray pool is exhausted (the warps that are left run at memory latency, not
//       at issue rate) the records of a leaf are prefetched when the leaf is entered and a pushed far child when
//       it is pushed, whatever bit 8 says
constexpr int kVarWideNode = 1, kVarSignMask = 2, kVarLeafPrefetch1 = 4, kVarNoPrefetch = 8, kVarTriBranchFree = 16,
              kVarSinglePop = 32, kVarChunkDecode = 64, kVarLeaf2 = 256, kVarLeaf4 = 512, kVarDrainPrefetch = 1024;

template <bool F32, int VAR>
__device__ __forceinline__ void prefetch_next(const SceneView &sc, uint32_t ref, uint32_t rc, bool draining = false) {
  if ((VAR & kVarDrainPrefetch) && draining) {
    if (rc == kBranch) prefetch_l1(sc.nodes + ref);
    else prefetch_leaf<F32>(sc.tris, ref, rc);
    return;
  }
  if (VAR & kVarNoPrefetch) return;
  if (rc == kBranch) {
    prefetch_l1(sc.nodes + ref);
  } else if (VAR & kVarLeafPrefetch1) {
    const size_t rec = F32 ? sizeof(TriRecordF32) : sizeof(TriRecordF64);
    prefetch_l1(reinterpret_cast<const char *>(sc.tris) + (size_t)ref * rec);
  } else {
    prefetch_leaf<F32>(sc.tris, ref, rc);
  }
}

struct NodeWords { // one PairNode as loaded
  double b[2][6];
  uint32_t ref0, ref1, cnt0, cnt1, axis;
};

template <bool WIDE> __device__ __forceinline__ NodeWords load_pair_node(const PairNode *n) {
  NodeWords w;
  if (WIDE) {
    const char *p = reinterpret_cast<const char *>(n);
    double a0, a1, a2, a3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p + 32 * k));
      (&w.b[0][0])[4 * k + 0] = a0, (&w.b[0][0])[4 * k + 1] = a1, (&w.b[0][0])[4 * k + 2] = a2, (&w.b[0][0])[4 * k + 3] = a3;
    }
    uint32_t m0, m1, m2, m3, m4, m5, m6, m7;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3), "=r"(m4), "=r"(m5), "=r"(m6), "=r"(m7)
                 : "l"(p + 96));
    w.ref0 = m0, w.ref1 = m1, w.cnt0 = m2, w.cnt1 = m3, w.axis = m4;
  } else {
    const double2 *np = reinterpret_cast<const double2 *>(n);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double2 v = __ldg(np + k);
      (&w.b[0][0])[2 * k] = v.x, (&w.b[0][0])[2 * k + 1] = v.y;
    }
    const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(np + 6));
    w.ref0 = meta.x, w.ref1 = meta.y, w.cnt0 = meta.z, w.cnt1 = meta.w;
    w.axis = __ldg(reinterpret_cast<const uint32_t *>(np + 7));
  }
  return w;
}

template <class IO, bool F32, int S, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int POLICY, unsigned CHUNK, int VAR>
__device__ __forceinline__ void trace_state_machine(const SceneView &sc, const IO &io, unsigned long long n,
                                                    unsigned long long *work, TravStack<S, CAP> &st,
                                                    unsigned long long *gcounters) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;

  RayD r;
  double hit_t = 0.0, tmax_any = 0.0;
  uint32_t ref = 0, rc = kIdle, item = 0;
  int sp = 0;
  uint32_t pool_next = 0, pool_end = 0;
  bool exhausted = false;
  uint32_t iter = 0, born = 0; // warp iterations so far / at the time this lane's ray was loaded
  typename IO::Chunk chunk;
  TravCounters cnt = {0u, 0u, 0u};
  unsigned int nrays = 0;
  r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = 0.0;
  r.sgn = 0u;

  for (;; iter++) {
    // ---- A. refill idle lanes from the warp's pool of ray indices ----------------------------------
    const unsigned idle = __ballot_sync(kFullMask, rc == kIdle);
    if (idle) {
      if (!exhausted && (__popc(idle) >= REFILL_MIN || idle == kFullMask)) {
        if (pool_next == pool_end) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(work, (unsigned long long)CHUNK);
          base = __shfl_sync(kFullMask, base, 0);
          if (base >= n) {
            exhausted = true;
          } else {
            pool_next = (uint32_t)base;
            pool_end = (uint32_t)((base + CHUNK < n) ? base + CHUNK : n);
            if (VAR & kVarChunkDecode) io.prepare(chunk, pool_next);
          }
        }
        if (!exhausted) {
          const unsigned avail = pool_end - pool_next, want = __popc(idle);
          const unsigned rank = __popc(idle & lt_mask);
          if (rc == kIdle && rank < avail) {
            item = pool_next + rank;
            double ox, oy, oz, dx, dy, dz, t0;
            if ((VAR & kVarChunkDecode) ? io.load(chunk, item, ox, oy, oz, dx, dy, dz, t0)
                                        : io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
              ray_setup(r, ox, oy, oz, dx, dy, dz);
              hit_t = t0;
              if (ANYHIT) tmax_any = t0;
              sp = 0;
              if (IO::kTracksCost) born = iter;
              if (COUNT) nrays++;
              bool enter = false;
              if (!sc.empty) {
                double tm;
                if (COUNT) cnt.nodes++;
                enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4],
                                  sc.root_box[5], r, hit_t, tm);
              }
              if (enter && sc.root_cnt != 0u) {
                ref = sc.root_ref, rc = sc.root_cnt;
                if (COUNT && rc != kBranch) cnt.tris += rc;
              } else {
                io.finish(item, false);
              }
            }
          }
          pool_next += (want < avail) ? want : avail;
        }
      } else if (exhausted && idle == kFullMask) {
        break;
      }
    }

    // ---- B. which step bodies run this iteration ------------------------------------------------------
    const bool at_inner = (rc == kBranch);
    const bool at_leaf = (rc - 1u) < (kIdle - 1u); // 1 <= rc < kIdle
    bool run_inner = true, run_leaf = true;
    if (POLICY == 0) {
      const unsigned m_inner = __ballot_sync(kFullMask, at_inner);
      const unsigned m_leaf = __ballot_sync(kFullMask, at_leaf);
      if (!(m_inner | m_leaf)) continue;
      run_inner = __popc(m_inner) > __popc(m_leaf);
      run_leaf = !run_inner;
    } else if (POLICY >= 3) {
      // threshold vote (A/B): a body runs only when at least POLICY * 2 lanes wait for it, unless the
      // other body does not qualify either (then the fuller one runs); the skipped lanes wait one iteration
      const int n_inner = __popc(__ballot_sync(kFullMask, at_inner));
      const int n_leaf = __popc(__ballot_sync(kFullMask, at_leaf));
      constexpr int T = POLICY * 2;
      run_inner = n_inner >= T;
      run_leaf = n_leaf >= T;
      if (!run_inner && !run_leaf) {
        run_inner = n_inner >= n_leaf;
        run_leaf = !run_inner;
      }
    }

    if (run_inner && at_inner) {
      // ---- INNER: one 128-byte PairNode, both children tested (equivalence: traverse.cuh) -------------
      const NodeWords nw = load_pair_node<(VAR & kVarWideNode) != 0>(sc.nodes + ref);
      double t0, t1;
      const bool h0 = slab_test(nw.b[0][0], nw.b[0][1], nw.b[0][2], nw.b[0][3], nw.b[0][4], nw.b[0][5], r, hit_t, t0);
      const bool h1 = slab_test(nw.b[1][0], nw.b[1][1], nw.b[1][2], nw.b[1][3], nw.b[1][4], nw.b[1][5], r, hit_t, t1);
      if (COUNT) cnt.nodes += 2;
      const bool sgn = ((r.sgn >> nw.axis) & 1u) != 0u; // dirSign[axis]
      if (h0 && h1) { // near = data[dirSign[axis]] first, far pushed with its tmin (bvh_accel.cc:818-823)
        st.put(sp++, sgn ? t0 : t1, sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1);
        if ((VAR & kVarDrainPrefetch) && exhausted && (sgn ? nw.cnt0 : nw.cnt1) != 0u)
          prefetch_next<F32, VAR>(sc, sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1, true);
        if (COUNT) cnt.max_stack = max(cnt.max_stack, (unsigned int)sp + 1u);
        ref = sgn ? nw.ref1 : nw.ref0, rc = sgn ? nw.cnt1 : nw.cnt0;
      } else if (h0) {
        ref = nw.ref0, rc = nw.cnt0;
      } else if (h1) {
        ref = nw.ref1, rc = nw.cnt1;
      } else {
        rc = 0u;
      }
      if (rc != 0u) {
        if (COUNT && rc != kBranch) cnt.tris += rc;
        prefetch_next<F32, VAR>(sc, ref, rc, exhausted);
      }
    } else if (run_leaf && at_leaf) {
      // ---- LEAF: one triangle of TestLeafNode (bvh_accel.cc:640-697), in indices_ order -----------------
      constexpr int kPerStep = (VAR & kVarLeaf4) ? 4 : ((VAR & kVarLeaf2) ? 2 : 1);
#pragma unroll
      for (int k = 0; k < kPerStep; k++) {
        const TriEdges tv = load_tri_edges<F32>(sc.tris, ref);
        double u, v;
        if (tri_test_edges<(VAR & kVarTriBranchFree) != 0>(hit_t, u, v, tv, r)) {
          io.accept(item, hit_t, u, v, tv.face, tv.mat);
          if (ANYHIT && hit_t < tmax_any) { // occluded: closest-hit Traverse would return t < tmax
            io.finish(item, true);
            if (IO::kTracksCost && iter - born > io.m.hot_steps) io.mark_hot(item);
            rc = kIdle;
          }
        }
        if (rc == kIdle) break;
        ref++;
        rc--;
        if (rc == 0u) break;
      }
    }

    // ---- C. pop: the reference's pop-time (tmin <= hitT) decision; empty stack = ray finished ----------
    if (rc == 0u) {
      if (VAR & kVarSinglePop) {
        if (sp == 0) {
          io.finish(item, false);
          rc = kIdle;
        } else {
          double tm;
          st.get(--sp, tm, ref, rc);
          if (tm <= hit_t && rc != 0u) {
            if (COUNT && rc != kBranch) cnt.tris += rc;
            prefetch_next<F32, VAR>(sc, ref, rc);
          } else {
            rc = 0u; // culled at pop time: pops again next iteration
          }
        }
      } else {
        for (;;) {
          if (sp == 0) {
            io.finish(item, false);
            if (IO::kTracksCost && iter - born > io.m.hot_steps) io.mark_hot(item);
            rc = kIdle;
            break;
          }
          double tm;
          st.get(--sp, tm, ref, rc);
          if (tm <= hit_t && rc != 0u) {
            if (COUNT && rc != kBranch) cnt.tris += rc;
            prefetch_next<F32, VAR>(sc, ref, rc);
            break;
          }
          rc = 0u;
        }
      }
    }
  }

  if (COUNT) {
    unsigned long long a = cnt.nodes, b = cnt.tris, c = nrays;
    unsigned int m = cnt.max_stack;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(kFullMask, a, o);
      b += __shfl_down_sync(kFullMask, b, o);
      c += __shfl_down_sync(kFullMask, c, o);
      m = max(m, __shfl_down_sync(kFullMask, m, o));
    }
    if (lane == 0) {
      atomicAdd(&gcounters[0], a);
      atomicAdd(&gcounters[1], b);
      atomicAdd(&gcounters[2], c);
      atomicMax(&gcounters[3], (unsigned long long)m);
    }
  }
}

} // namespace mb200

#endif

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

// K2 over a caller's ray buffer (mb200_trace_closest).
struct IOClosest {
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
};

// K1 fused into K2: the camera ray of work item i (one jittered sample of one pixel) is generated in
// the refill step -- Camera::GenerateRay (camera.cc:222-240) after PathTrace's jitter (render.cc:386-393)
// -- and never stored; hits[i] receives the 32-byte record.
struct IOCamera {
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
    camera_sample(p, x, y, pass, rng, dx, dy, dz);
    ox = p.frame.origin[0], oy = p.frame.origin[1], oz = p.frame.origin[2];
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
  const QRay *q;
  mb200_hit *hits;
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
  const QRay *q;
  float *contrib; // [items]
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
};

// ---- the state machine ------------------------------------------------------------------------------
// REFILL_MIN: idle lanes that trigger a refill; CHUNK: ray indices taken from the global counter per
// atomicAdd (32 keeps the end-of-launch imbalance small: a launch of 2 M rays is only ~18 per lane).
template <class IO, bool F32, int S, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int POLICY, unsigned CHUNK>
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
  TravCounters cnt = {0u, 0u, 0u};
  unsigned int nrays = 0;
  r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = 0.0;
  r.sx = r.sy = r.sz = false;

  for (;;) {
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
          }
        }
        if (!exhausted) {
          const unsigned avail = pool_end - pool_next, want = __popc(idle);
          const unsigned rank = __popc(idle & lt_mask);
          if (rc == kIdle && rank < avail) {
            item = pool_next + rank;
            double ox, oy, oz, dx, dy, dz, t0;
            if (io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
              ray_setup(r, ox, oy, oz, dx, dy, dz);
              hit_t = t0;
              if (ANYHIT) tmax_any = t0;
              sp = 0;
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
    }

    if (run_inner && at_inner) {
      // ---- INNER: one 128-byte PairNode, both children tested (equivalence: traverse.cuh) -------------
      const double2 *np = reinterpret_cast<const double2 *>(sc.nodes + ref);
      const double2 a0 = __ldg(np + 0), a1 = __ldg(np + 1), a2 = __ldg(np + 2);
      const double2 b0 = __ldg(np + 3), b1 = __ldg(np + 4), b2 = __ldg(np + 5);
      const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(np + 6));
      const uint32_t axis = __ldg(reinterpret_cast<const uint32_t *>(np + 7));
      double t0, t1;
      const bool h0 = slab_test(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, r, hit_t, t0);
      const bool h1 = slab_test(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, r, hit_t, t1);
      if (COUNT) cnt.nodes += 2;
      const bool sgn = (axis == 0) ? r.sx : ((axis == 1) ? r.sy : r.sz);
      if (h0 && h1) { // near = data[dirSign[axis]] first, far pushed with its tmin (bvh_accel.cc:818-823)
        st.put(sp++, sgn ? t0 : t1, sgn ? meta.x : meta.y, sgn ? meta.z : meta.w);
        if (COUNT) cnt.max_stack = max(cnt.max_stack, (unsigned int)sp + 1u);
        ref = sgn ? meta.y : meta.x, rc = sgn ? meta.w : meta.z;
      } else if (h0) {
        ref = meta.x, rc = meta.z;
      } else if (h1) {
        ref = meta.y, rc = meta.w;
      } else {
        rc = 0u;
      }
      if (rc != 0u) {
        if (rc == kBranch) {
          prefetch_l1(sc.nodes + ref);
        } else {
          if (COUNT) cnt.tris += rc;
          prefetch_leaf<F32>(sc.tris, ref, rc);
        }
      }
    } else if (run_leaf && at_leaf) {
      // ---- LEAF: one triangle of TestLeafNode (bvh_accel.cc:640-697), in indices_ order -----------------
      const TriEdges tv = load_tri_edges<F32>(sc.tris, ref);
      double u, v;
      if (tri_test_edges(hit_t, u, v, tv, r)) {
        io.accept(item, hit_t, u, v, tv.face, tv.mat);
        if (ANYHIT && hit_t < tmax_any) { // occluded: closest-hit Traverse would return t < tmax
          io.finish(item, true);
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
          io.finish(item, false);
          rc = kIdle;
          break;
        }
        double tm;
        st.get(--sp, tm, ref, rc);
        if (tm <= hit_t && rc != 0u) {
          if (rc == kBranch) {
            prefetch_l1(sc.nodes + ref);
          } else {
            if (COUNT) cnt.tris += rc;
            prefetch_leaf<F32>(sc.tris, ref, rc);
          }
          break;
        }
        rc = 0u;
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

// trace_mr.cuh -- the traversal state machine with K rays per lane (K2 closest hit / K4 any hit).
//
// Why.  In the one-ray-per-lane machine (trace_sm.cuh) every warp iteration issues both step bodies --
// INNER: one PairNode visit (two slab tests), LEAF: one triangle test -- and each lane uses exactly one
// of them, so each body runs with about half of the lanes whatever the scheduling policy (measured:
// 17 and 13 of 32 lanes, profiles/r1_trace_camera_sm_lines.md; vote / threshold policies did not help,
// DESIGN.md §5).  The only way past 50 % is to give a lane useful work in BOTH bodies of an iteration.
// Here every lane owns K ray slots.  Per iteration it advances one of its rays that is at an inner node
// and one of its rays that is at a leaf; with K = 4 a lane almost always has both kinds.
//
// Where the state lives.  A ray between steps is 96 bytes in shared memory (six 16-byte units in a
// column layout: unit u of slot s of thread t at smem[(s*6+u)*blockDim + t], so a warp's LDS.128 /
// STS.128 are conflict-free):
//     u0 ox oy | u1 oz hitT | u2 ix iy | u3 iz dx | u4 dy dz | u5 ref, rc, sp + dirSign bits, item
// Registers hold only three K-bit masks (which slots are at an inner node / at a leaf / need a pop), so
// the two bodies can use the whole register file for their own temporaries and their global loads are
// issued back to back before either body's arithmetic.  The traversal stacks are per slot in L1-cached
// local memory (measured equal to the shared-memory stack).
//
// Exactness.  Per ray the sequence of node visits, triangle tests and pop-time culling decisions is
// unchanged (the step bodies are those of trace_sm.cuh / traverse.cuh); only the interleaving between
// rays differs, and rays do not interact.  Counters (COUNT) equal the oracle's.
#ifndef MALLIE_B200_TRACE_MR_CUH_
#define MALLIE_B200_TRACE_MR_CUH_

#include "trace_sm.cuh"

namespace mb200 {

constexpr int kSlotUnits = 6; // 16-byte units per ray slot

// S: stack entries per slot kept in shared memory right behind the slot's six units
template <int K, int S = 0> struct SlotMem {
  uint4 *base; // this thread's column
  int stride;  // blockDim.x
  __device__ __forceinline__ uint4 *unit(int s, int u) const { return base + (s * (kSlotUnits + S) + u) * stride; }
  __device__ __forceinline__ double2 ld2(int s, int u) const {
    const uint4 v = *unit(s, u);
    return make_double2(__longlong_as_double((long long)(((unsigned long long)v.y << 32) | v.x)),
                        __longlong_as_double((long long)(((unsigned long long)v.w << 32) | v.z)));
  }
  __device__ __forceinline__ void st2(int s, int u, double a, double b) const {
    const unsigned long long x = (unsigned long long)__double_as_longlong(a), y = (unsigned long long)__double_as_longlong(b);
    *unit(s, u) = make_uint4((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)y, (uint32_t)(y >> 32));
  }
  // hitT is the second double of unit 1
  __device__ __forceinline__ void st_hit_t(int s, double t) const {
    reinterpret_cast<double *>(unit(s, 1))[1] = t;
  }
};

// K rays per lane.  CAP: stack entries per ray (local memory).  The IO policies are those of trace_sm.cuh;
// any-hit sources must provide tmax_of(item) (the occlusion distance, re-read on the rare accepted hit).
template <class IO, bool F32, int K, int S, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, unsigned CHUNK>
__device__ __forceinline__ void trace_multi_ray(const SceneView &sc, const IO &io, unsigned long long n,
                                                unsigned long long *work, const SlotMem<K, S> sm,
                                                unsigned long long *gcounters) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr uint32_t kAll = (1u << K) - 1u;

  // (tmin, ref, cnt) stack entries per slot: the first S in shared memory, the rest in L1-cached local memory
  uint4 stk[K * ((CAP > S) ? (CAP - S) : 1)];
  auto stk_put = [&](int s, uint32_t k, const uint4 &e) {
    if (S > 0 && k < (uint32_t)S) *sm.unit(s, kSlotUnits + (int)k) = e;
    else stk[s * (CAP - S) + (int)k - S] = e;
  };
  auto stk_get = [&](int s, uint32_t k) -> uint4 {
    return (S > 0 && k < (uint32_t)S) ? *sm.unit(s, kSlotUnits + (int)k) : stk[s * (CAP - S) + (int)k - S];
  };
  uint32_t m_inner = 0u, m_leaf = 0u, m_pop = 0u;
  uint32_t pool_next = 0, pool_end = 0;
  bool exhausted = false;
  TravCounters cnt = {0u, 0u, 0u};
  unsigned int nrays = 0;

  for (;;) {
    // ---- A. refill: one idle slot per lane, from the warp's pool of consecutive ray indices -------------
    const uint32_t busy = m_inner | m_leaf | m_pop;
    const unsigned idle = __ballot_sync(kFullMask, busy != kAll);
    if (idle) {
      const unsigned none_busy = __ballot_sync(kFullMask, busy == 0u);
      if (!exhausted && (__popc(idle) >= REFILL_MIN || none_busy == kFullMask)) {
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
          if (busy != kAll && rank < avail) {
            const uint32_t item = pool_next + rank;
            double ox, oy, oz, dx, dy, dz, t0;
            if (io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
              RayD r;
              ray_setup(r, ox, oy, oz, dx, dy, dz);
              if (COUNT) nrays++;
              bool enter = false;
              if (!sc.empty) {
                double tm;
                if (COUNT) cnt.nodes++;
                enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4],
                                  sc.root_box[5], r, t0, tm);
              }
              if (enter && sc.root_cnt != 0u) {
                const int s = __ffs((int)(~busy & kAll)) - 1;
                sm.st2(s, 0, ox, oy);
                sm.st2(s, 1, oz, t0);
                sm.st2(s, 2, r.ix, r.iy);
                sm.st2(s, 3, r.iz, dx);
                sm.st2(s, 4, dy, dz);
                *sm.unit(s, 5) = make_uint4(sc.root_ref, sc.root_cnt, r.sgn << 16, item);
                if (sc.root_cnt == kBranch) {
                  m_inner |= 1u << s;
                } else {
                  m_leaf |= 1u << s;
                  if (COUNT) cnt.tris += sc.root_cnt;
                }
              } else {
                io.finish(item, false);
              }
            }
          }
          pool_next += (want < avail) ? want : avail;
        }
      } else if (exhausted && none_busy == kFullMask) {
        break;
      }
    }

    // ---- B. pick one slot per body, fetch their state and issue both global loads -----------------------
    const bool do_inner = m_inner != 0u, do_leaf = m_leaf != 0u;
    const int si = do_inner ? __ffs((int)m_inner) - 1 : 0;
    const int sl = do_leaf ? __ffs((int)m_leaf) - 1 : 0;

    RayD ri;
    double hit_i = 0.0;
    uint4 wi = make_uint4(0u, 0u, 0u, 0u);
    NodeWords nw;
    if (do_inner) {
      const double2 a = sm.ld2(si, 0), b = sm.ld2(si, 1), c = sm.ld2(si, 2), d = sm.ld2(si, 3);
      wi = *sm.unit(si, 5);
      ri.ox = a.x, ri.oy = a.y, ri.oz = b.x, hit_i = b.y, ri.ix = c.x, ri.iy = c.y, ri.iz = d.x;
      ri.dx = ri.dy = ri.dz = 0.0;
      ri.sgn = (wi.z >> 16) & 7u;
      nw = load_pair_node<true>(sc.nodes + wi.x);
    }
    RayD rl;
    double hit_l = 0.0;
    uint4 wl = make_uint4(0u, 0u, 0u, 0u);
    TriEdges tv;
    if (do_leaf) {
      const double2 a = sm.ld2(sl, 0), b = sm.ld2(sl, 1), d = sm.ld2(sl, 3), e = sm.ld2(sl, 4);
      wl = *sm.unit(sl, 5);
      rl.ox = a.x, rl.oy = a.y, rl.oz = b.x, hit_l = b.y, rl.dx = d.y, rl.dy = e.x, rl.dz = e.y;
      rl.ix = rl.iy = rl.iz = 0.0;
      rl.sgn = 0u;
      tv = load_tri_edges<F32>(sc.tris, wl.x);
    }

    // ---- C. INNER: one 128-byte PairNode, both children tested (equivalence: traverse.cuh) --------------
    if (do_inner) {
      double t0, t1;
      const bool h0 = slab_test(nw.b[0][0], nw.b[0][1], nw.b[0][2], nw.b[0][3], nw.b[0][4], nw.b[0][5], ri, hit_i, t0);
      const bool h1 = slab_test(nw.b[1][0], nw.b[1][1], nw.b[1][2], nw.b[1][3], nw.b[1][4], nw.b[1][5], ri, hit_i, t1);
      if (COUNT) cnt.nodes += 2;
      const bool sgn = (((wi.z >> 16) >> nw.axis) & 1u) != 0u;
      uint32_t ref, rc;
      uint32_t sp = wi.z & 0xFFFFu;
      if (h0 && h1) { // near = data[dirSign[axis]] first, far pushed with its tmin (bvh_accel.cc:818-823)
        const double tf = sgn ? t0 : t1;
        const unsigned long long tb = (unsigned long long)__double_as_longlong(tf);
        stk_put(si, sp, make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1));
        sp++;
        if (COUNT) cnt.max_stack = max(cnt.max_stack, sp + 1u);
        ref = sgn ? nw.ref1 : nw.ref0, rc = sgn ? nw.cnt1 : nw.cnt0;
      } else if (h0) {
        ref = nw.ref0, rc = nw.cnt0;
      } else if (h1) {
        ref = nw.ref1, rc = nw.cnt1;
      } else {
        ref = 0u, rc = 0u;
      }
      *sm.unit(si, 5) = make_uint4(ref, rc, (wi.z & 0xFFFF0000u) | sp, wi.w);
      if (rc != kBranch) {
        m_inner &= ~(1u << si);
        if (rc != 0u) {
          m_leaf |= 1u << si;
          if (COUNT) cnt.tris += rc;
        } else {
          m_pop |= 1u << si;
        }
      }
    }

    // ---- D. LEAF: one triangle of TestLeafNode (bvh_accel.cc:640-697), in indices_ order -----------------
    if (do_leaf) {
      double u, v;
      bool done = false;
      if (tri_test_edges<false>(hit_l, u, v, tv, rl)) {
        io.accept(wl.w, hit_l, u, v, tv.face, tv.mat);
        sm.st_hit_t(sl, hit_l);
        if constexpr (ANYHIT) {
          if (hit_l < io.tmax_of(wl.w)) { // occluded: closest-hit Traverse would return t < tmax
            io.finish(wl.w, true);
            done = true;
          }
        }
      }
      if (done) {
        m_leaf &= ~(1u << sl);
      } else {
        const uint32_t rc = wl.y - 1u;
        *sm.unit(sl, 5) = make_uint4(wl.x + 1u, rc, wl.z, wl.w);
        if (rc == 0u) {
          m_leaf &= ~(1u << sl);
          m_pop |= 1u << sl;
        }
      }
    }

    // ---- E. pop: the reference's pop-time (tmin <= hitT) decision; empty stack = ray finished -------------
    while (m_pop) {
      const int s = __ffs((int)m_pop) - 1;
      m_pop &= ~(1u << s);
      const uint4 w = *sm.unit(s, 5);
      const double hit_t = sm.ld2(s, 1).y;
      uint32_t sp = w.z & 0xFFFFu, ref = 0u, rc = 0u;
      for (;;) {
        if (sp == 0u) break;
        const uint4 e = stk_get(s, --sp);
        const double tm = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
        if (tm <= hit_t && e.w != 0u) {
          ref = e.z, rc = e.w;
          break;
        }
      }
      if (rc == 0u) {
        io.finish(w.w, false);
      } else {
        *sm.unit(s, 5) = make_uint4(ref, rc, (w.z & 0xFFFF0000u) | sp, w.w);
        if (rc == kBranch) {
          m_inner |= 1u << s;
        } else {
          m_leaf |= 1u << s;
          if (COUNT) cnt.tris += rc;
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

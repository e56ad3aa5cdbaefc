// trace_tr.cuh -- the traversal state machine with TWO rays per lane and ONE step body per warp iteration.
//
// Why.  In trace_sm.cuh every warp iteration issues both step bodies -- INNER (one pair-node visit) and LEAF (one
// triangle test) -- and each lane uses exactly one of them, so each body runs with about half of the lanes
// (17 and 13 of 32, profiles/r2_trace_camera.md) and the kernel, which is bound by issue slots, pays ~8 issued
// instructions per lane-step where a phase-coherent warp would pay ~4.  Letting the warp vote for one body per
// iteration does not help with one ray per lane (the other half of the lanes then waits; round 1 measured +5 %).
// Here every lane owns two rays.  Per iteration the warp runs the body more of its lanes can take part in, and a
// lane takes part if EITHER of its rays is in that phase: with the phases split ~55 / 45 that is ~80 % of the lanes
// instead of ~50 %.
//
// Where the state lives.  Round 1's K-rays-per-lane variant kept all ray state in shared memory and reloaded it on
// every step; its bookkeeping ate the gain.  Here a ray's mutable state is six registers (hitT, ref, rc, item and
// a word packing stack depth | direction signs | birth iteration) and BOTH rays' mutable state stays in registers:
// an "active" set the bodies work on and a "parked" set, exchanged by plain register moves when the lane's other ray
// is the one that can take the step.  Only the read-only part of a ray (origin, direction, 1/direction: 72 bytes, written
// once when the ray is loaded) lives in shared memory, in the conflict-free column layout of the old traversal stack,
// and the registers hold a copy of the half a body needs (origin + 1/direction for INNER, origin + direction for LEAF),
// refreshed only when the lane changes ray or phase.  The traversal stacks (one per ray) are in L1-cached local memory,
// which measured equal to shared memory for the one-ray machine.  Register budget and occupancy are the
// one-ray machine's (64 registers, 8 CTAs of 128 threads per SM).
//
// Exactness.  Per ray the sequence of node visits, triangle tests and pop-time culling decisions is unchanged (the step
// bodies are those of trace_sm.cuh / traverse.cuh); only the interleaving between rays differs, and rays do not
// interact.  Counters (COUNT) equal the oracle's.
#ifndef MALLIE_B200_TRACE_TR_CUH_
#define MALLIE_B200_TRACE_TR_CUH_

#include "trace_sm.cuh"

namespace mb200 {

constexpr uint32_t kSpMask = 0x3FFu; // bits 0-9 of the packed word: stack depth; 10-12: direction signs; 13-31: birth iteration
constexpr int kFatUnits = 5;         // 16-byte units per ray: (ox, oy) (oz, ix) (iy, iz) (dx, dy) (dz, tmax)

__device__ __forceinline__ void sts128(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void lds128m(uint32_t addr, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
}

template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, int REFILL_MIN, int HYST, unsigned CHUNK>
__device__ __forceinline__ void trace_two_ray_machine(const SceneView &sc, const IO &io, unsigned long long n,
                                                      unsigned long long *work, uint32_t fat_base, uint32_t unit_stride,
                                                      unsigned long long *gcounters) {
  static_assert(!IO::kFused, "the fused frame form uses the one-ray machine");
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint4 stk[2 * CAP]; // local memory: CAP entries per ray

  // mutable state of the lane's two rays: the active set (the bodies work on it) and the parked set
  double a_t = 0.0, p_t = 0.0;
  uint32_t a_ref = 0, a_rc = kIdle, a_item = 0, a_sb = 0;
  uint32_t p_ref = 0, p_rc = kIdle, p_item = 0, p_sb = 0;
  uint32_t act = 0; // index (0 / 1) of the ray in the active set: selects its shared-memory units and its stack
  // register copy of the read-only half the current body needs
  double ox = 0.0, oy = 0.0, oz = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
  uint32_t fat_tag = 0xFFu; // ray index | phase << 1 of what the copy holds; 0xFF: nothing
  uint32_t pool_next = 0, pool_end = 0;
  bool exhausted = false;
  uint32_t iter = 0;
  int phase = 0; // the body the warp runs: 0 INNER, 1 LEAF
  TravCounters cnt = {0u, 0u, 0u};
  unsigned int nrays = 0;

  auto unit = [&](uint32_t ray, int k) -> uint32_t { return fat_base + (ray * kFatUnits + (uint32_t)k) * unit_stride; };

  for (;; iter++) {
    // ---- A. refill: one new ray per lane that has an empty place (invariant: active empty => parked empty) ----------
    const unsigned want_mask = __ballot_sync(kFullMask, p_rc == kIdle);
    if (want_mask) {
      const unsigned none_mask = __ballot_sync(kFullMask, a_rc == kIdle);
      if (!exhausted && (__popc(want_mask) >= REFILL_MIN || none_mask == kFullMask)) {
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
          const unsigned avail = pool_end - pool_next, want = __popc(want_mask);
          const unsigned rank = __popc(want_mask & lt_mask);
          if (p_rc == kIdle && rank < avail) {
            const uint32_t item = pool_next + rank;
            double lx, ly, lz, dx, dy, dz, t0;
            if (io.load(item, lx, ly, lz, dx, dy, dz, t0)) {
              RayD r;
              ray_setup(r, lx, ly, lz, dx, dy, dz);
              const bool into_active = a_rc == kIdle;
              const uint32_t rix = into_active ? act : (act ^ 1u);
              sts128(unit(rix, 0), r.ox, r.oy);
              sts128(unit(rix, 1), r.oz, r.ix);
              sts128(unit(rix, 2), r.iy, r.iz);
              sts128(unit(rix, 3), r.dx, r.dy);
              sts128(unit(rix, 4), r.dz, t0);
              if ((fat_tag & 1u) == rix) fat_tag = 0xFFu; // that ray's units changed under the register copy
              if (COUNT) nrays++;
              uint32_t nref = 0, nrc = kIdle;
              bool enter = false;
              if (!sc.empty) {
                double tm;
                if (COUNT) cnt.nodes++;
                enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4],
                                  sc.root_box[5], r, DBL_MAX, tm);
              }
              if (enter && sc.root_cnt != 0u) {
                nref = sc.root_ref, nrc = sc.root_cnt;
                if (COUNT && nrc != kBranch) cnt.tris += nrc;
              } else {
                io.finish(item, false);
              }
              const uint32_t nsb = (r.sgn << 10) | (iter << 13);
              if (nrc != kIdle) {
                if (into_active) a_t = DBL_MAX, a_ref = nref, a_rc = nrc, a_item = item, a_sb = nsb;
                else p_t = DBL_MAX, p_ref = nref, p_rc = nrc, p_item = item, p_sb = nsb;
              }
            }
          }
          pool_next += (want < avail) ? want : avail;
        }
      }
      if (exhausted && none_mask == kFullMask) {
        if (__ballot_sync(kFullMask, a_rc == kIdle) == kFullMask) break; // (a ray loaded above would clear its bit)
      }
    }

    // ---- B. which body: the one more lanes can take part in with either of their rays (with hysteresis) ---------------
    const bool a_in = a_rc == kBranch, p_in = p_rc == kBranch;
    const bool a_lf = (a_rc - 1u) < (kShade - 1u), p_lf = (p_rc - 1u) < (kShade - 1u);
    const int n_in = __popc(__ballot_sync(kFullMask, a_in | p_in));
    const int n_lf = __popc(__ballot_sync(kFullMask, a_lf | p_lf));
    if ((n_in | n_lf) == 0) continue;
    if (phase == 0) {
      if (n_lf > n_in + HYST || n_in == 0) phase = 1;
    } else {
      if (n_in > n_lf + HYST || n_lf == 0) phase = 0;
    }

    // ---- C. the lane's other ray is the one that can step: exchange the register sets -------------------------------
    if (phase ? (!a_lf && p_lf) : (!a_in && p_in)) {
      const double tt = a_t;
      a_t = p_t, p_t = tt;
      uint32_t w;
      w = a_ref, a_ref = p_ref, p_ref = w;
      w = a_rc, a_rc = p_rc, p_rc = w;
      w = a_item, a_item = p_item, p_item = w;
      w = a_sb, a_sb = p_sb, p_sb = w;
      act ^= 1u;
    }

    if (phase == 0) {
      if (a_rc == kBranch) {
        // ---- INNER: one 128-byte PairNode, both children tested (equivalence: traverse.cuh) -----------------------
        if (fat_tag != act) { // origin + 1 / direction of the active ray
          lds128m(unit(act, 0), ox, oy);
          lds128m(unit(act, 1), oz, gx);
          lds128m(unit(act, 2), gy, gz);
          fat_tag = act;
        }
        RayD r;
        r.ox = ox, r.oy = oy, r.oz = oz, r.ix = gx, r.iy = gy, r.iz = gz;
        r.dx = r.dy = r.dz = 0.0;
        r.sgn = (a_sb >> 10) & 7u;
        const NodeWords nw = load_pair_node(sc.nodes + a_ref);
        double t0, t1;
        const bool h0 = slab_test(nw.b[0][0], nw.b[0][1], nw.b[0][2], nw.b[0][3], nw.b[0][4], nw.b[0][5], r, a_t, t0);
        const bool h1 = slab_test(nw.b[1][0], nw.b[1][1], nw.b[1][2], nw.b[1][3], nw.b[1][4], nw.b[1][5], r, a_t, t1);
        if (COUNT) cnt.nodes += 2;
        const bool sgn = ((r.sgn >> nw.axis) & 1u) != 0u; // dirSign[axis]
        if (h0 && h1) { // near = data[dirSign[axis]] first, far pushed with its tmin (bvh_accel.cc:818-823)
          const double tf = sgn ? t0 : t1;
          const unsigned long long tb = (unsigned long long)__double_as_longlong(tf);
          stk[act * CAP + (a_sb & kSpMask)] = make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1);
          a_sb++;
          if (COUNT) cnt.max_stack = max(cnt.max_stack, (a_sb & kSpMask) + 1u);
          a_ref = sgn ? nw.ref1 : nw.ref0, a_rc = sgn ? nw.cnt1 : nw.cnt0;
        } else if (h0) {
          a_ref = nw.ref0, a_rc = nw.cnt0;
        } else if (h1) {
          a_ref = nw.ref1, a_rc = nw.cnt1;
        } else {
          a_rc = 0u;
        }
        if (COUNT && a_rc != 0u && a_rc != kBranch) cnt.tris += a_rc;
      }
    } else {
      if ((a_rc - 1u) < (kShade - 1u)) {
        // ---- LEAF: one triangle of TestLeafNode (bvh_accel.cc:640-697), in indices_ order ---------------------------
        if (fat_tag != (act | 2u)) { // origin + direction of the active ray
          double skip;
          lds128m(unit(act, 0), ox, oy);
          lds128m(unit(act, 1), oz, skip);
          lds128m(unit(act, 3), gx, gy);
          lds128m(unit(act, 4), gz, skip);
          fat_tag = act | 2u;
        }
        RayD r;
        r.ox = ox, r.oy = oy, r.oz = oz, r.dx = gx, r.dy = gy, r.dz = gz;
        r.ix = r.iy = r.iz = 0.0;
        r.sgn = 0u;
        const TriEdges tv = load_tri_edges<TRI>(sc.trav_tris, a_ref);
        double u, v;
        if (tri_test_edges(a_t, u, v, tv, r)) {
          io.accept(a_item, a_t, u, v, tv.face, tv.mat);
          if (ANYHIT) {
            double dz_, tmax;
            lds128m(unit(act, 4), dz_, tmax);
            if (a_t < tmax) { // occluded: closest-hit Traverse would return t < tmax
              io.finish(a_item, true);
              if (IO::kTracksCost && ((iter - (a_sb >> 13)) & 0x7FFFFu) > io.m.hot_steps) io.mark_hot(a_item);
              a_rc = kIdle;
            }
          }
        }
        if (a_rc != kIdle) {
          a_ref++;
          a_rc--;
        }
      }
    }

    // ---- D. pop: the reference's pop-time (tmin <= hitT) decision; empty stack = ray finished ------------------------
    if (a_rc == 0u) {
      for (;;) {
        if ((a_sb & kSpMask) == 0u) {
          if (IO::kTracksCost && ((iter - (a_sb >> 13)) & 0x7FFFFu) > io.m.hot_steps) io.mark_hot(a_item);
          io.finish(a_item, false);
          a_rc = kIdle;
          break;
        }
        a_sb--;
        const uint4 e = stk[act * CAP + (a_sb & kSpMask)];
        const double tm = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
        a_ref = e.z, a_rc = e.w;
        if (tm <= a_t && a_rc != 0u) {
          if (COUNT && a_rc != kBranch) cnt.tris += a_rc;
          break;
        }
        a_rc = 0u;
      }
    }
    // the active place emptied: the parked ray (if any) moves up, so that "active empty => parked empty" holds
    if (a_rc == kIdle && p_rc != kIdle) {
      a_t = p_t, a_ref = p_ref, a_rc = p_rc, a_item = p_item, a_sb = p_sb;
      p_rc = kIdle;
      act ^= 1u;
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

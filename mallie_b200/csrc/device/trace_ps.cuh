// trace_ps.cuh -- DEVELOPMENT VARIANT: rays move between the warps of a CTA so that a warp runs ONE step body at a time
// with (nearly) all of its lanes.
//
// Why.  In trace_sm.cuh a ray stays on its lane for life, every warp iteration issues both step bodies and each runs
// with ~13.5 of 32 lanes; the FP64 pipe charges a warp instruction the same whatever its active lanes
// (profiles/r2_fp64_halfwarp_probe.txt), so lanes per instruction is the one lever left.  Variants that
// keep the ray on its lane (vote, two rays per lane, one slot per phase) did not pay.  Here a ray lives in a SLOT of
// the CTA's shared memory and is in one of two queues, INNER (next step: a pair-node visit) or LEAF (next step: a
// triangle test).  A warp takes up to 32 rays of ONE queue, loads what that step body needs into registers, runs only
// that body until fewer than kStay of its lanes are still in the phase, writes the few words that changed back to the
// slots and files every ray under its new phase (or retires it and frees the slot).  New rays enter through the same
// loop: 32 free slots at a time are filled from the global work counter.
//
// Slot layout (16-byte units, unit u of slot s at  base + (u * NS + s) * 16):
//   0 (ox, oy)  1 (oz, ix)  2 (iy, iz)  3 (dx, dy)  4 (dz, -)     written once when the ray is loaded
//   5 (hitT, ref | rc << 32)            6 (item, sp, -, -)        the state a round changes
//   7 .. 7 + S - 1                      traversal stack entries; deeper ones in a slot-indexed global scratch
// Queues are three LIFO arrays of slot ids (INNER, LEAF, free) under one CTA-wide spin lock.  In this first version every
// queue operation is a critical section executed by LANE 0 ALONE (the other lanes hand their slot ids over through a
// 32-word staging area per warp); `held` counts the rays that are in registers, for the exit test.
//
// Measured (profiles/r2_trace_phase_sorted_variant.md): bit-identical, FP64 instructions per launch -33 ... -37 % -- the
// bodies do run a third fuller -- and 32.0 ms per bench frame against 19.9, because the serial critical sections add
// ~40 % instructions at one active lane, 8-14 x the shared-memory wavefronts and lock waits at 24 warps per SM.  A
// version with warp-wide critical sections (one queue entry per lane) corrupted its queues under some interleavings
// although the same protocol passes a standalone stress test; DESIGN.md section 10 lists what the next version needs.
//
// Exactness: per ray the sequence of pair-node visits, triangle tests and pop-time decisions is trace_sm.cuh's (the
// bodies are the same code); only where and when a step executes changes.  Needs the octant copies of the pair nodes.
#ifndef MALLIE_B200_TRACE_PS_CUH_
#define MALLIE_B200_TRACE_PS_CUH_

#include "trace_sm.cuh"

namespace mb200 {

constexpr int kPsRoUnits = 5, kPsMutUnits = 2;
constexpr unsigned kPsStageBytes = 8 * 32 * 4; // up to 8 warps per CTA

struct PsShared { // control words at the start of the CTA's dynamic shared memory
  unsigned int lock, n_inner, n_leaf, n_free, held, exhausted, pad0, pad1;
};

__device__ __forceinline__ void ps_sts128(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ps_lds128(uint32_t addr, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ps_sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ps_lds128u(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

// NS slots per CTA, S stack entries per slot in shared memory, CAP stack entries in all, STAY: a round ends when fewer
// lanes than this are still in its phase.
template <class IO, int TRI, int NS, int S, int CAP, bool ANYHIT, int STAY, unsigned CHUNK>
__device__ __forceinline__ void trace_phase_sorted(const SceneView &sc, const IO &io, unsigned long long n,
                                                   unsigned long long *work, unsigned char *smem_raw, uint4 *ovf_cta) {
  static_assert(!IO::kFused, "wavefront sources only");
  const unsigned lane = threadIdx.x & 31u;
  PsShared *ctl = reinterpret_cast<PsShared *>(smem_raw);
  unsigned short *q_inner = reinterpret_cast<unsigned short *>(smem_raw + sizeof(PsShared));
  unsigned short *q_leaf = q_inner + NS;
  unsigned short *q_free = q_leaf + NS;
  // per-warp staging of 32 words: the queue operations are done by lane 0 alone, the other lanes hand their slot ids over here
  volatile unsigned int *stage = reinterpret_cast<volatile unsigned int *>(smem_raw + ((sizeof(PsShared) + 3 * NS * 2 + 15) & ~15u)) +
                                 (threadIdx.x >> 5) * 32u;
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)(((sizeof(PsShared) + 3 * NS * 2 + 15) & ~15u) + kPsStageBytes);
  auto unit = [&](uint32_t slot, int u) -> uint32_t { return slots + ((uint32_t)u * NS + slot) * 16u; };

  if (threadIdx.x == 0) {
    ctl->lock = 0u, ctl->n_inner = 0u, ctl->n_leaf = 0u, ctl->n_free = NS, ctl->held = 0u, ctl->exhausted = 0u, ctl->pad0 = 0u, ctl->pad1 = 0u;
  }
  for (unsigned i = threadIdx.x; i < NS; i += blockDim.x) q_free[i] = (unsigned short)i;
  __syncthreads();

  // files the warp's rays: lane 0 alone works on the queues (under the lock); where = 0 INNER, 1 LEAF, 2 free, 3 nothing
  auto file = [&](uint32_t slot, int where, unsigned released) {
    stage[lane] = (slot << 2) | (uint32_t)where;
    __syncwarp();
    if (lane == 0) {
      while (atomicCAS(&ctl->lock, 0u, 1u) != 0u) __nanosleep(20);
      __threadfence_block();
      volatile PsShared *c = ctl;
      unsigned ni = c->n_inner, nl = c->n_leaf, nf = c->n_free;
      for (int k = 0; k < 32; k++) {
        const unsigned e = stage[k];
        const unsigned w = e & 3u;
        if (w == 0u) reinterpret_cast<volatile unsigned short *>(q_inner)[ni++] = (unsigned short)(e >> 2);
        else if (w == 1u) reinterpret_cast<volatile unsigned short *>(q_leaf)[nl++] = (unsigned short)(e >> 2);
        else if (w == 2u) reinterpret_cast<volatile unsigned short *>(q_free)[nf++] = (unsigned short)(e >> 2);
      }
      c->n_inner = ni, c->n_leaf = nl, c->n_free = nf;
      c->held = c->held - released;
      __threadfence_block();
      atomicExch(&ctl->lock, 0u);
    }
    __syncwarp();
  };

  for (unsigned guard = 0;; guard++) {
    // ---- what to do next -----------------------------------------------------------------------------------------
    int job = 0; // 0 wait, 1 INNER round, 2 LEAF round, 3 refill, 4 exit
    unsigned take = 0;
    uint32_t slot = 0;
    if (lane == 0) {
      while (atomicCAS(&ctl->lock, 0u, 1u) != 0u) __nanosleep(20);
      __threadfence_block();
      volatile PsShared *c = ctl;
      const unsigned ci = c->n_inner, cl = c->n_leaf, cf = c->n_free, held = c->held;
      const bool can_refill = !c->exhausted && cf > 0u;
      const unsigned big = ci > cl ? ci : cl;
      unsigned j = 0, t = 0;
      if (big >= 32u || (!can_refill && big > 0u)) {
        j = (cl >= ci) ? 2u : 1u;
        t = big < 32u ? big : 32u;
        volatile unsigned short *q = reinterpret_cast<volatile unsigned short *>((j == 2u) ? q_leaf : q_inner);
        for (unsigned k = 0; k < t; k++) stage[k] = q[big - t + k];
        if (j == 2u) c->n_leaf = big - t;
        else c->n_inner = big - t;
        c->held = held + t;
      } else if (can_refill) {
        j = 3u;
        t = cf < 32u ? cf : 32u;
        for (unsigned k = 0; k < t; k++) stage[k] = reinterpret_cast<volatile unsigned short *>(q_free)[cf - t + k];
        c->n_free = cf - t;
        c->held = held + t;
      } else if (held == 0u) {
        j = 4u;
      }
      __threadfence_block();
      atomicExch(&ctl->lock, 0u);
      job = (int)j, take = t;
    }
    __syncwarp();
    job = __shfl_sync(kFullMask, job, 0);
    take = __shfl_sync(kFullMask, take, 0);
    if (lane < take) slot = stage[lane];
    __syncwarp();
    if (job == 4) break;
    if (job == 0) {
      __nanosleep(200);
      continue;
    }
    const bool mine = lane < take;

    if (job == 3) {
      // ---- refill: `take` new rays into free slots ------------------------------------------------------------
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(work, (unsigned long long)take);
      base = __shfl_sync(kFullMask, base, 0);
      if (base + take >= n && lane == 0) *reinterpret_cast<volatile unsigned int *>(&ctl->exhausted) = 1u;
      int where = mine ? 2 : 3; // a slot that gets no ray goes back to the free list
      if (mine && base + lane < n) {
        const uint32_t item = (uint32_t)(base + lane);
        double ox, oy, oz, dx, dy, dz, t0;
        if (io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
          RayD r;
          ray_setup(r, ox, oy, oz, dx, dy, dz);
          bool enter = false;
          if (!sc.empty) {
            double tm;
            enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], r,
                              DBL_MAX, tm);
          }
          if (enter && sc.root_cnt != 0u) {
            uint32_t ref = sc.root_ref;
            const uint32_t rc = sc.root_cnt;
            if (rc == kBranch) ref += r.sgn * sc.num_pair_nodes; // into the octant's copy
            ps_sts128(unit(slot, 0), r.ox, r.oy);
            ps_sts128(unit(slot, 1), r.oz, r.ix);
            ps_sts128(unit(slot, 2), r.iy, r.iz);
            ps_sts128(unit(slot, 3), r.dx, r.dy);
            ps_sts128(unit(slot, 4), r.dz, 0.0);
            const unsigned long long tb = (unsigned long long)__double_as_longlong(DBL_MAX);
            ps_sts128u(unit(slot, 5), (uint32_t)tb, (uint32_t)(tb >> 32), ref, rc);
            ps_sts128u(unit(slot, 6), item, 0u, 0u, 0u);
            where = (rc == kBranch) ? 0 : 1;
          } else {
            io.finish(item, false);
          }
        }
      }
      file(slot, where, take);
      continue;
    }

    // ---- a round of one phase -------------------------------------------------------------------------------------
    double hit_t = 0.0;
    uint32_t ref = 0, rc = kIdle, item = 0, spw = 0;
    if (mine) {
      uint32_t a, b, c, d;
      ps_lds128u(unit(slot, 5), a, b, ref, rc);
      hit_t = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
      ps_lds128u(unit(slot, 6), item, spw, c, d);
    }
    int sp = (int)spw;
    uint4 *ovf = ovf_cta + (size_t)slot * (CAP - S);
    auto put = [&](int k, double tmin, uint32_t pref, uint32_t pcnt) {
      const unsigned long long tb = (unsigned long long)__double_as_longlong(tmin);
      if (k < S) ps_sts128u(unit(slot, kPsRoUnits + kPsMutUnits + k), (uint32_t)tb, (uint32_t)(tb >> 32), pref, pcnt);
      else ovf[k - S] = make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), pref, pcnt);
    };
    auto get = [&](int k, double &tmin, uint32_t &pref, uint32_t &pcnt) {
      uint4 e;
      if (k < S) ps_lds128u(unit(slot, kPsRoUnits + kPsMutUnits + k), e.x, e.y, e.z, e.w);
      else e = ovf[k - S];
      tmin = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
      pref = e.z, pcnt = e.w;
    };
    // pop: the reference's pop-time (tmin <= hitT) decision; empty stack = ray finished
    auto pop = [&]() {
      for (;;) {
        if (sp == 0) {
          io.finish(item, false);
          rc = kIdle;
          break;
        }
        double tm;
        get(--sp, tm, ref, rc);
        if (tm <= hit_t && rc != 0u) break;
        rc = 0u;
      }
    };

    if (job == 1) { // INNER: origin and 1 / direction in registers
      RayD r;
      r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0;
      r.dx = r.dy = r.dz = 0.0, r.sgn = 0u;
      if (mine) {
        ps_lds128(unit(slot, 0), r.ox, r.oy);
        ps_lds128(unit(slot, 1), r.oz, r.ix);
        ps_lds128(unit(slot, 2), r.iy, r.iz);
      }
      for (unsigned it = 0;; it++) {
        const bool at_inner = (rc == kBranch);
        const int staying = __popc(__ballot_sync(kFullMask, at_inner));
        if (staying == 0 || (it > 0 && staying < STAY)) break;
        if (at_inner) {
          const NodeWords nw = load_pair_node<false>(sc.nodes_oct + ref);
          double t0, t1;
          const bool h0 = slab_test_oct<0>(nw.b[0], r, hit_t, t0);
          const bool h1 = slab_test_oct<0>(nw.b[1], r, hit_t, t1);
          if (h0 && h1) {
            if (sp < CAP) put(sp, t1, nw.ref1, nw.cnt1);
            sp++;
            ref = nw.ref0, rc = nw.cnt0;
          } else if (h0) {
            ref = nw.ref0, rc = nw.cnt0;
          } else if (h1) {
            ref = nw.ref1, rc = nw.cnt1;
          } else {
            rc = 0u;
          }
          if (rc == 0u) pop();
        }
      }
    } else { // LEAF: origin and direction in registers
      RayD r;
      r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = 0.0;
      r.ix = r.iy = r.iz = 0.0, r.sgn = 0u;
      if (mine) {
        double unused;
        ps_lds128(unit(slot, 0), r.ox, r.oy);
        ps_lds128(unit(slot, 1), r.oz, unused);
        ps_lds128(unit(slot, 3), r.dx, r.dy);
        ps_lds128(unit(slot, 4), r.dz, unused);
      }
      for (unsigned it = 0;; it++) {
        const bool at_leaf = (rc - 1u) < (kShade - 1u);
        const int staying = __popc(__ballot_sync(kFullMask, at_leaf));
        if (staying == 0 || (it > 0 && staying < STAY)) break;
        if (at_leaf) {
          double u, v;
          const TriEdges tv = load_tri_edges<TRI>(sc.trav_tris, ref);
          if (tri_test_edges(hit_t, u, v, tv, r)) {
            io.accept(item, hit_t, u, v, tv.face, tv.mat);
            if constexpr (ANYHIT) {
              if (hit_t < io.tmax_of(item)) {
                io.finish(item, true);
                rc = kIdle;
              }
            }
          }
          if (rc != kIdle) {
            ref++, rc--;
            if (rc == 0u) pop();
          }
        }
      }
    }

    // ---- write back what changed, file the rays under their new phase ---------------------------------------------
    int where = 3;
    if (mine) {
      if (rc == kIdle) {
        where = 2;
      } else {
        const unsigned long long tb = (unsigned long long)__double_as_longlong(hit_t);
        ps_sts128u(unit(slot, 5), (uint32_t)tb, (uint32_t)(tb >> 32), ref, rc);
        ps_sts128u(unit(slot, 6), item, (uint32_t)sp, 0u, 0u);
        where = (rc == kBranch) ? 0 : 1;
      }
    }
    file(slot, where, take);
    if (guard > 400000000u) __trap(); // development guard against a scheduling bug
  }
}

} // namespace mb200

#endif

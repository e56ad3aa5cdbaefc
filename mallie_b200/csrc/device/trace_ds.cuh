// trace_ds.cuh -- the traversal state machine with two ray SLOTS per lane, one per phase ("dual slot").
//
// Why.  In trace_sm.cuh a lane owns one ray, every warp iteration issues both step bodies -- INNER (one pair-node visit)
// and LEAF (one triangle test) -- and the lane uses the one its ray is in, so each body runs with about half of the lanes
// (13.5 of 32 on average) in a kernel that is bound by issue slots.  trace_tr.cuh gave every lane two rays and let the
// warp vote for one body per iteration; lane use rose to 23 of 32, but choosing the ray, exchanging register sets and
// reloading the ray's read-only half at low width cost more instructions than the fuller bodies saved.
//
// Here a lane has an I slot and an L slot, each with its OWN registers: the I slot holds a ray that is at an inner node
// (origin, 1/direction, hitT, node), the L slot a ray that is inside a leaf (origin, direction, hitT, triangle range).
// Both bodies run every iteration, each on its own fixed registers -- no per-step selects, no votes -- and a lane whose two
// rays are in different phases takes part in both.  A ray changes slot when it changes phase: I -> L when it reaches a leaf
// and the L slot is free, L -> I when its leaf is done, its stack yields an inner node and the I slot is free; when both
// want to change, they swap.  A ray that cannot move waits in place (its body skips it).  The read-only half a slot needs
// (1/direction or direction, plus the origin) comes from the ray's 80 bytes in shared memory, written once when the ray is
// loaded; the mutable state (hitT, ref, count, item, stack depth) moves by register.  Each ray has its own stack in
// L1-cached local memory, addressed by the ray's index (0 / 1), which travels with it.
//
// Exactness.  Per ray the sequence of node visits, triangle tests and pop-time culling decisions is unchanged (the step
// bodies are those of trace_sm.cuh / traverse.cuh); only the interleaving between rays differs, and rays do not interact.
#ifndef MALLIE_B200_TRACE_DS_CUH_
#define MALLIE_B200_TRACE_DS_CUH_

#include "trace_tr.cuh"

namespace mb200 {

// packed word of a ray: bits 0-9 stack depth, bit 10 ray index (which stack / which shared-memory units), 11-31 birth iteration
constexpr uint32_t kDsRayBit = 0x400u;

// OCT: the scene has octant copies of its pair nodes (layout.h): the I slot then needs no direction signs.
template <class IO, int TRI, int CAP, bool ANYHIT, bool COUNT, bool OCT, int REFILL_MIN, unsigned CHUNK>
__device__ __forceinline__ void trace_dual_slot_machine(const SceneView &sc, const IO &io, unsigned long long n,
                                                        unsigned long long *work, uint32_t fat_base, uint32_t unit_stride,
                                                        unsigned long long *gcounters) {
  static_assert(!IO::kFused, "the fused frame form uses the one-ray machine");
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint4 stk[2 * CAP]; // local memory: CAP entries per ray index

  // I slot: a ray at an inner node (i_rc == kBranch), or waiting to enter a leaf (1 <= i_rc < kShade), or none (kIdle)
  double i_ox = 0.0, i_oy = 0.0, i_oz = 0.0, i_ix = 0.0, i_iy = 0.0, i_iz = 0.0, i_t = 0.0;
  uint32_t i_ref = 0, i_rc = kIdle, i_item = 0, i_sb = 0, i_sgn = 0;
  // L slot: a ray inside a leaf (1 <= l_rc < kShade), or waiting to visit an inner node (kBranch), or none (kIdle)
  double l_ox = 0.0, l_oy = 0.0, l_oz = 0.0, l_dx = 0.0, l_dy = 0.0, l_dz = 0.0, l_t = 0.0;
  uint32_t l_ref = 0, l_rc = kIdle, l_item = 0, l_sb = 0;

  uint32_t pool_next = 0, pool_end = 0;
  bool exhausted = false;
  uint32_t iter = 0;
  TravCounters cnt = {0u, 0u, 0u};
  unsigned int nrays = 0;

  auto unit = [&](uint32_t ray, int k) -> uint32_t { return fat_base + (ray * kFatUnits + (uint32_t)k) * unit_stride; };
  auto is_leaf = [](uint32_t rc) -> bool { return (rc - 1u) < (kShade - 1u); };
  // pops ray `sb`'s stack until an entry survives the reference's pop-time test; false: the stack is empty
  auto pop = [&](double hit_t, uint32_t &sb, uint32_t &ref, uint32_t &rc) -> bool {
    for (;;) {
      if ((sb & kSpMask) == 0u) return false;
      sb--;
      const uint4 e = stk[((sb & kDsRayBit) ? CAP : 0) + (sb & kSpMask)];
      const double tm = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
      if (tm <= hit_t && e.w != 0u) {
        ref = e.z, rc = e.w;
        if (COUNT && rc != kBranch) cnt.tris += rc;
        return true;
      }
    }
  };

  for (;; iter++) {
    // ---- A. refill: a new ray for every lane whose I slot is empty ------------------------------------------------------
    const unsigned want_mask = __ballot_sync(kFullMask, i_rc == kIdle);
    if (want_mask) {
      const unsigned none_mask = __ballot_sync(kFullMask, i_rc == kIdle && l_rc == kIdle);
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
          if (i_rc == kIdle && rank < avail) {
            const uint32_t item = pool_next + rank;
            double ox, oy, oz, dx, dy, dz, t0;
            if (io.load(item, ox, oy, oz, dx, dy, dz, t0)) {
              RayD r;
              ray_setup(r, ox, oy, oz, dx, dy, dz);
              // the ray index the lane's other ray does not use
              const uint32_t rix = (l_rc != kIdle && !(l_sb & kDsRayBit)) ? 1u : 0u;
              sts128(unit(rix, 0), r.ox, r.oy);
              sts128(unit(rix, 1), r.oz, r.ix);
              sts128(unit(rix, 2), r.iy, r.iz);
              sts128(unit(rix, 3), r.dx, r.dy);
              sts128(unit(rix, 4), r.dz, t0);
              if (COUNT) nrays++;
              bool enter = false;
              if (!sc.empty) {
                double tm;
                if (COUNT) cnt.nodes++;
                enter = slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4],
                                  sc.root_box[5], r, DBL_MAX, tm);
              }
              if (enter && sc.root_cnt != 0u) {
                i_ox = r.ox, i_oy = r.oy, i_oz = r.oz, i_ix = r.ix, i_iy = r.iy, i_iz = r.iz;
                i_t = DBL_MAX, i_item = item, i_sgn = r.sgn;
                i_sb = (rix ? kDsRayBit : 0u) | (iter << 11);
                i_ref = sc.root_ref, i_rc = sc.root_cnt;
                if (sc.root_cnt == kBranch) {
                  if (OCT) i_ref += r.sgn * sc.num_pair_nodes; // into the octant's copy (layout.h)
                } else if (COUNT) {
                  cnt.tris += i_rc;
                }
              } else {
                io.finish(item, false);
              }
            }
          }
          pool_next += (want < avail) ? want : avail;
        }
      }
      if (exhausted && none_mask == kFullMask) break;
    }

    // ---- B. INNER: one 128-byte PairNode of the I slot's ray, both children tested (equivalence: traverse.cuh) ----------
    if (i_rc == kBranch) {
      RayD r;
      r.ox = i_ox, r.oy = i_oy, r.oz = i_oz, r.ix = i_ix, r.iy = i_iy, r.iz = i_iz;
      r.dx = r.dy = r.dz = 0.0;
      r.sgn = i_sgn;
      double t0, t1;
      bool h0, h1, sgn;
      NodeWords nw;
      if (OCT) { // octant copies: boxes and children pre-ordered for the ray's octant
        nw = load_pair_node<false>(sc.nodes_oct + i_ref);
        h0 = slab_test_oct<0>(nw.b[0], r, i_t, t0);
        h1 = slab_test_oct<0>(nw.b[1], r, i_t, t1);
        sgn = false;
      } else {
        nw = load_pair_node(sc.nodes + i_ref);
        h0 = slab_test(nw.b[0][0], nw.b[0][1], nw.b[0][2], nw.b[0][3], nw.b[0][4], nw.b[0][5], r, i_t, t0);
        h1 = slab_test(nw.b[1][0], nw.b[1][1], nw.b[1][2], nw.b[1][3], nw.b[1][4], nw.b[1][5], r, i_t, t1);
        sgn = ((r.sgn >> nw.axis) & 1u) != 0u;
      }
      if (COUNT) cnt.nodes += 2;
      if (h0 && h1) { // near first, far pushed with its tmin (bvh_accel.cc:818-823)
        const double tf = sgn ? t0 : t1;
        const unsigned long long tb = (unsigned long long)__double_as_longlong(tf);
        stk[((i_sb & kDsRayBit) ? CAP : 0) + (i_sb & kSpMask)] =
            make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), sgn ? nw.ref0 : nw.ref1, sgn ? nw.cnt0 : nw.cnt1);
        i_sb++;
        if (COUNT) cnt.max_stack = max(cnt.max_stack, (i_sb & kSpMask) + 1u);
        i_ref = sgn ? nw.ref1 : nw.ref0, i_rc = sgn ? nw.cnt1 : nw.cnt0;
      } else if (h0) {
        i_ref = nw.ref0, i_rc = nw.cnt0;
      } else if (h1) {
        i_ref = nw.ref1, i_rc = nw.cnt1;
      } else {
        i_rc = 0u;
      }
      if (COUNT && i_rc != 0u && i_rc != kBranch) cnt.tris += i_rc;
      if (i_rc == 0u && !pop(i_t, i_sb, i_ref, i_rc)) { // nothing left: the ray is finished
        if (IO::kTracksCost && ((iter - (i_sb >> 11)) & 0x1FFFFFu) > io.m.hot_steps) io.mark_hot(i_item);
        io.finish(i_item, false);
        i_rc = kIdle;
      }
    }

    // ---- C. LEAF: one triangle of the L slot's ray (TestLeafNode, bvh_accel.cc:640-697, in indices_ order) ----------------
    if (is_leaf(l_rc)) {
      RayD r;
      r.ox = l_ox, r.oy = l_oy, r.oz = l_oz, r.dx = l_dx, r.dy = l_dy, r.dz = l_dz;
      r.ix = r.iy = r.iz = 0.0;
      r.sgn = 0u;
      const TriEdges tv = load_tri_edges<TRI>(sc.trav_tris, l_ref);
      double u, v;
      bool stop = false;
      if (tri_test_edges(l_t, u, v, tv, r)) {
        io.accept(l_item, l_t, u, v, tv.face, tv.mat);
        if constexpr (ANYHIT) {
          if (l_t < io.tmax_of(l_item)) { // occluded: closest-hit Traverse would return t < tmax
            io.finish(l_item, true);
            stop = true;
          }
        }
      }
      if (!stop) {
        l_ref++;
        l_rc--;
        if (l_rc == 0u && !pop(l_t, l_sb, l_ref, l_rc)) { // leaf done and nothing left
          io.finish(l_item, false);
          stop = true;
        }
      }
      if (stop) {
        if (IO::kTracksCost && ((iter - (l_sb >> 11)) & 0x1FFFFFu) > io.m.hot_steps) io.mark_hot(l_item);
        l_rc = kIdle;
      }
    }

    // ---- D. rays that changed phase change slot (when the other slot is free, or by swapping) -------------------------------
    const bool i_wants_l = is_leaf(i_rc), l_wants_i = l_rc == kBranch;
    if (i_wants_l | l_wants_i) {
      if (i_wants_l && (l_wants_i || l_rc == kIdle)) {
        // I -> L (the L ray, if any, goes the other way below): mutable state by register, origin + direction from shared memory
        const double nt = i_t;
        const uint32_t nref = i_ref, nrc = i_rc, nitem = i_item, nsb = i_sb;
        if (l_wants_i) { // swap: the L ray takes the I slot
          const uint32_t rix = (l_sb & kDsRayBit) ? 1u : 0u;
          lds128m(unit(rix, 0), i_ox, i_oy);
          lds128m(unit(rix, 1), i_oz, i_ix);
          lds128m(unit(rix, 2), i_iy, i_iz);
          i_t = l_t, i_ref = l_ref, i_rc = kBranch, i_item = l_item, i_sb = l_sb;
          if (!OCT) { // dirSign (bvh_accel.cc:787-790) re-derived from the direction
            double dx, dy, dz, w;
            lds128m(unit(rix, 3), dx, dy);
            lds128m(unit(rix, 4), dz, w);
            i_sgn = (dx < 0.0 ? 1u : 0u) | (dy < 0.0 ? 2u : 0u) | (dz < 0.0 ? 4u : 0u);
          }
        } else {
          i_rc = kIdle;
        }
        const uint32_t rix = (nsb & kDsRayBit) ? 1u : 0u;
        double w;
        lds128m(unit(rix, 0), l_ox, l_oy);
        lds128m(unit(rix, 1), l_oz, w);
        lds128m(unit(rix, 3), l_dx, l_dy);
        lds128m(unit(rix, 4), l_dz, w);
        l_t = nt, l_ref = nref, l_rc = nrc, l_item = nitem, l_sb = nsb;
      } else if (l_wants_i && i_rc == kIdle) {
        // L -> I
        const uint32_t rix = (l_sb & kDsRayBit) ? 1u : 0u;
        lds128m(unit(rix, 0), i_ox, i_oy);
        lds128m(unit(rix, 1), i_oz, i_ix);
        lds128m(unit(rix, 2), i_iy, i_iz);
        if (!OCT) {
          double dx, dy, dz, w;
          lds128m(unit(rix, 3), dx, dy);
          lds128m(unit(rix, 4), dz, w);
          i_sgn = (dx < 0.0 ? 1u : 0u) | (dy < 0.0 ? 2u : 0u) | (dz < 0.0 ? 4u : 0u);
        }
        i_t = l_t, i_ref = l_ref, i_rc = kBranch, i_item = l_item, i_sb = l_sb;
        l_rc = kIdle;
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

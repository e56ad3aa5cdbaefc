// bvh_build_gpu.cu -- BVHAccel::Build (bvh_accel.cc:36-482) on the device, bit-identical tree.
//
// The reference builder is a depth-first recursion; its result, however, is a pure function of the
// triangle set of every node, so the same tree can be grown LEVEL BY LEVEL with all nodes of a level
// processed side by side (host/bvh_build.cc is the sequential statement of the same algorithm; tests hold
// this file against it, against the oracle's builder and against the reference's tree fingerprints, bit
// for bit; tests/lsbuild_model.py is a numpy model of the steps below):
//
//   per level, over all active segments [l, r) of the index array:
//     bounds      min / max of the cached triangle bounds, +- kEPS            (bvh_accel.cc:285-315)
//     leaf test   n < minLeafPrimitives || depth >= maxTreeDepth               (bvh_accel.cc:341)
//     histogram   64 bins x 3 axes of triangle-bound minima / maxima           (bvh_accel.cc:82-142)
//     sweep       63 SAH planes per axis, evaluation order of the reference    (bvh_accel.cc:156-255)
//     partition   std::partition on (p0+p1+p2)[axis] < 3*pos                   (bvh_accel.cc:257-277,402)
//     children    [l, mid), [mid, r); object-median fallback                   (bvh_accel.cc:405-430)
//   afterwards: subtree sizes bottom-up, pre-order numbers top-down (left child = parent + 1), emit.
//
// The partition is the only order-sensitive step.  libstdc++'s bidirectional std::partition swaps the
// k-th misplaced element from the left with the k-th misplaced element from the RIGHT and touches nothing
// else, so the final arrangement follows from two prefix sums over the "misplaced" flags; no sequential
// pass is needed.  All floating-point expressions are those of host/bvh_build.cc (-fmad=false).
//
// Segmented reductions take a block-wide fast path when a whole CTA lies inside one segment (always true
// near the root, where contention on per-segment atomics would otherwise serialise) and fall back to
// per-element atomics elsewhere (deep levels: many small segments, no contention).
//
// scene_build_device (mb200_scene_build) continues on the device: branch ranks number the 128-byte pair
// nodes, both children's boxes are written into the parent and the leaf-ordered triangle records are
// emitted -- what scene.cc::relayout_bvh does on the host, byte for byte -- so that only the mesh crosses
// PCIe on the way in and nothing has to come back.
#include <cuda_runtime.h>

#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../host/bvh_build.h"
#include "layout.h"
#include "mallie_b200.h"
#include "scene.h"

namespace mb200 {

namespace {

constexpr double kPad = DBL_EPSILON * 1024.0;
constexpr uint32_t kNoSeg = 0xFFFFFFFFu;
constexpr int kThreads = 256;

#define CUB(call)                                   \
  do {                                              \
    cudaError_t e_ = (call);                        \
    if (e_ != cudaSuccess) return e_;               \
  } while (0)

// order-preserving map double <-> uint64 (for atomicMin / atomicMax on doubles)
__device__ __forceinline__ unsigned long long dkey(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

struct Segs { // one level's active segments, sorted by l
  uint32_t *l, *r, *node;
};

struct SegWork { // per segment, per level
  unsigned long long *kmin; // [S*3] keys of min(lo)
  unsigned long long *kmax; // [S*3] keys of max(hi)
  double *bmin, *bmax;      // [S*3] padded bounds
  double *scale, *step;     // [S*3]
  uint8_t *leaf;            // [S]
  uint32_t *hslot;          // [S] histogram slot of a branch segment
  int32_t *axis;            // [S]
  double *thresh;           // [S] 3 * cut position
  uint32_t *ntrue;          // [S]
  uint32_t *isbranch;       // [S] 0 / 1 (scanned into child ordinals)
};

// ---- triangle cache ----------------------------------------------------------------------------------
__global__ void k_tri_cache(const double *__restrict__ v, const uint32_t *__restrict__ f, uint32_t nt,
                            double *__restrict__ lo, double *__restrict__ hi, double *__restrict__ cs,
                            uint32_t *__restrict__ idx) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const double *p0 = v + 3 * (size_t)f[3 * t + 0], *p1 = v + 3 * (size_t)f[3 * t + 1], *p2 = v + 3 * (size_t)f[3 * t + 2];
  for (int a = 0; a < 3; a++) {
    double mn = p0[a], mx = p0[a];
    if (p1[a] < mn) mn = p1[a];
    if (mx < p1[a]) mx = p1[a];
    if (p2[a] < mn) mn = p2[a];
    if (mx < p2[a]) mx = p2[a];
    lo[(size_t)a * nt + t] = mn;
    hi[(size_t)a * nt + t] = mx;
    cs[(size_t)a * nt + t] = p0[a] + p1[a] + p2[a];
  }
  idx[t] = t;
}

// ---- which segment does position i belong to (binary search over the sorted segment starts) ----------
__global__ void k_assign_sid(uint32_t n, const uint32_t *__restrict__ sl, const uint32_t *__restrict__ sr, uint32_t S,
                             uint32_t *__restrict__ sid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t lo = 0, hi = S; // last segment with l <= i
  while (hi - lo > 1) {
    const uint32_t m = (lo + hi) >> 1;
    if (sl[m] <= i) lo = m;
    else hi = m;
  }
  sid[i] = (S > 0 && sl[lo] <= i && i < sr[lo]) ? lo : kNoSeg;
}

__global__ void k_init_keys(uint32_t S3, unsigned long long *kmin, unsigned long long *kmax) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S3) kmin[i] = 0xFFFFFFFFFFFFFFFFull, kmax[i] = 0ull;
}

// ---- bounds -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_bounds(uint32_t n, uint32_t nt, const uint32_t *__restrict__ sid,
                                                    const uint32_t *__restrict__ idx, const double *__restrict__ lo,
                                                    const double *__restrict__ hi, unsigned long long *kmin,
                                                    unsigned long long *kmax) {
  __shared__ unsigned long long s_min[3][kThreads / 32], s_max[3][kThreads / 32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t first = blockIdx.x * blockDim.x, last = min(first + blockDim.x, n) - 1;
  const uint32_t s = i < n ? sid[i] : kNoSeg;
  const bool uniform = sid[first] != kNoSeg && sid[first] == sid[last]; // segments are contiguous
  unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0ull, 0ull, 0ull};
  if (s != kNoSeg) {
    const uint32_t t = idx[i];
    for (int a = 0; a < 3; a++) mn[a] = dkey(lo[(size_t)a * nt + t]), mx[a] = dkey(hi[(size_t)a * nt + t]);
  }
  if (!uniform) {
    if (s != kNoSeg)
      for (int a = 0; a < 3; a++) atomicMin(&kmin[3 * (size_t)s + a], mn[a]), atomicMax(&kmax[3 * (size_t)s + a], mx[a]);
    return;
  }
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = min(mn[a], __shfl_down_sync(0xFFFFFFFFu, mn[a], o));
      mx[a] = max(mx[a], __shfl_down_sync(0xFFFFFFFFu, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0) s_min[a][threadIdx.x >> 5] = mn[a], s_max[a][threadIdx.x >> 5] = mx[a];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    unsigned long long m0 = ~0ull, m1 = 0ull;
    for (int w = 0; w < kThreads / 32; w++) m0 = min(m0, s_min[a][w]), m1 = max(m1, s_max[a][w]);
    const uint32_t sg = sid[first];
    atomicMin(&kmin[3 * (size_t)sg + a], m0);
    atomicMax(&kmax[3 * (size_t)sg + a], m1);
  }
}

// ---- per segment: padded bounds, leaf decision, bin scale / step, histogram slot ----------------------
__global__ void k_seg_setup(uint32_t S, const uint32_t *__restrict__ sl, const uint32_t *__restrict__ sr, int level,
                            mb200_build_options opt, SegWork w, uint32_t *branch_counter) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t n = sr[s] - sl[s];
  const double nbins = (double)opt.bin_size;
  for (int a = 0; a < 3; a++) {
    const double bl = dunkey(w.kmin[3 * (size_t)s + a]) - kPad, bh = dunkey(w.kmax[3 * (size_t)s + a]) + kPad;
    w.bmin[3 * (size_t)s + a] = bl;
    w.bmax[3 * (size_t)s + a] = bh;
    const double extent = bh - bl;
    w.scale[3 * (size_t)s + a] = (extent > kPad) ? nbins / extent : 0.0;
    w.step[3 * (size_t)s + a] = extent * (1.0 / opt.bin_size);
  }
  const bool leaf = n < (uint32_t)opt.min_leaf_primitives || level >= opt.max_tree_depth;
  w.leaf[s] = leaf ? 1 : 0;
  w.isbranch[s] = leaf ? 0u : 1u;
  w.hslot[s] = leaf ? kNoSeg : atomicAdd(branch_counter, 1u);
  w.ntrue[s] = 0u;
}

// ---- histograms: hist[slot][axis][0 = minima, 1 = maxima][bin] ----------------------------------------
__global__ void __launch_bounds__(kThreads) k_histogram(uint32_t n, uint32_t nt, int nb, const uint32_t *__restrict__ sid,
                                                       const uint32_t *__restrict__ idx, const double *__restrict__ lo,
                                                       const double *__restrict__ hi, SegWork w, uint32_t *hist) {
  extern __shared__ uint32_t s_hist[]; // [6 * nb] when the block lies inside one segment
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t first = blockIdx.x * blockDim.x, last = min(first + blockDim.x, n) - 1;
  const uint32_t s0 = sid[first];
  const bool uniform = s0 != kNoSeg && s0 == sid[last] && nb <= 512;
  if (uniform && w.leaf[s0]) return;
  if (uniform) {
    for (int k = threadIdx.x; k < 6 * nb; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
  }
  const uint32_t s = i < n ? sid[i] : kNoSeg;
  if (s != kNoSeg && !w.leaf[s]) {
    const uint32_t t = idx[i];
    const double nbins = (double)nb;
    uint32_t *h = uniform ? s_hist : hist + (size_t)w.hslot[s] * 6 * nb;
    for (int a = 0; a < 3; a++) {
      const double bl = w.bmin[3 * (size_t)s + a], sc = w.scale[3 * (size_t)s + a];
      size_t qlo = (unsigned int)floor((lo[(size_t)a * nt + t] - bl) * sc);
      size_t qhi = (unsigned int)floor((hi[(size_t)a * nt + t] - bl) * sc);
      if ((double)qlo >= nbins) qlo = (size_t)nb - 1;
      if ((double)qhi >= nbins) qhi = (size_t)nb - 1;
      atomicAdd(&h[(a * 2 + 0) * nb + qlo], 1u);
      atomicAdd(&h[(a * 2 + 1) * nb + qhi], 1u);
    }
  }
  if (uniform) {
    __syncthreads();
    uint32_t *g = hist + (size_t)w.hslot[s0] * 6 * nb;
    for (int k = threadIdx.x; k < 6 * nb; k += blockDim.x)
      if (s_hist[k]) atomicAdd(&g[k], s_hist[k]);
  }
}

__device__ __forceinline__ double half_area2(const double lo[3], const double hi[3]) {
  const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return 2.0 * (dx * dy + dy * dz + dz * dx);
}

// ---- SAH sweep, one warp per branch segment (FindCutFromBinBuffer, bvh_accel.cc:156-255) ------------------
// The reference walks the nb - 1 candidate planes of an axis in order, carrying the running counts, and keeps the
// first plane whose cost is strictly below everything before it.  A plane's cost depends only on its own prefix
// counts, so the planes are evaluated 32 at a time (warp scan of the bin counts + carry) and the winner is the
// (cost, plane index) minimum -- lower index on equal cost, NaN costs never win, exactly as the sequential `<`.
__global__ void __launch_bounds__(kThreads) k_sweep(uint32_t S, const uint32_t *__restrict__ sl, const uint32_t *__restrict__ sr,
                                                   mb200_build_options opt, SegWork w, const uint32_t *__restrict__ hist) {
  const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (s >= S || w.leaf[s]) return; // warp-uniform
  const int nb = opt.bin_size;
  const uint32_t *h = hist + (size_t)w.hslot[s] * 6 * nb;
  double lo[3], hi[3];
  for (int a = 0; a < 3; a++) lo[a] = w.bmin[3 * (size_t)s + a], hi[a] = w.bmax[3 * (size_t)s + a];
  const size_t n = (size_t)sr[s] - sl[s];
  const double t_box = opt.cost_taabb, t_tri = 1.0 - opt.cost_taabb;
  const double whole = half_area2(lo, hi);
  const double inv_whole = (whole > kPad) ? 1.0 / whole : 0.0;
  double best_cost[3], best_pos[3];
  for (int a = 0; a < 3; a++) {
    const double step = w.step[3 * (size_t)s + a];
    double bc = DBL_MAX;
    int bi = -1; // no plane yet
    uint32_t carry_l = 0, carry_r = 0;
    double llo[3] = {lo[0], lo[1], lo[2]}, lhi[3] = {hi[0], hi[1], hi[2]};
    double rlo[3] = {lo[0], lo[1], lo[2]}, rhi[3] = {hi[0], hi[1], hi[2]};
    for (int base = 0; base < nb - 1; base += 32) {
      const int i = base + (int)lane;
      const bool act = i < nb - 1;
      uint32_t cl = act ? h[(a * 2 + 0) * nb + i] : 0u, cr = act ? h[(a * 2 + 1) * nb + i] : 0u;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t yl = __shfl_up_sync(0xFFFFFFFFu, cl, o), yr = __shfl_up_sync(0xFFFFFFFFu, cr, o);
        if ((int)lane >= o) cl += yl, cr += yr;
      }
      const size_t nl = (size_t)carry_l + cl, nr = n - ((size_t)carry_r + cr);
      carry_l += __shfl_sync(0xFFFFFFFFu, cl, 31);
      carry_r += __shfl_sync(0xFFFFFFFFu, cr, 31);
      if (act) {
        const double pos = lo[a] + (i + 0.5) * step;
        lhi[a] = pos;
        rlo[a] = pos;
        const double al = half_area2(llo, lhi), ar = half_area2(rlo, rhi);
        const double cost = 2.0f * t_box + (al * inv_whole) * (double)(nl)*t_tri + (ar * inv_whole) * (double)(nr)*t_tri;
        if (cost < bc) bc = cost, bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xFFFFFFFFu, bc, o);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      // a lane without a plane (bi < 0) holds DBL_MAX and every real plane is strictly below it
      if (oi >= 0 && (bi < 0 || oc < bc || (oc == bc && oi < bi))) bc = oc, bi = oi;
    }
    best_cost[a] = bc;
    best_pos[a] = lo[a] + ((bi < 0 ? 0 : bi) + 0.5) * step; // no plane: the first plane's position (bvh_accel.cc:170)
  }
  if (lane != 0) return;
  int axis = 0;
  double c = best_cost[0];
  if (c > best_cost[1]) axis = 1, c = best_cost[1];
  if (c > best_cost[2]) axis = 2, c = best_cost[2];
  w.axis[s] = axis;
  w.thresh[s] = best_pos[axis] * 3.0;
}

// ---- partition predicate + number of "left" triangles per segment --------------------------------------
__global__ void __launch_bounds__(kThreads) k_pred(uint32_t n, uint32_t nt, const uint32_t *__restrict__ sid,
                                                  const uint32_t *__restrict__ idx, const double *__restrict__ cs,
                                                  SegWork w, uint8_t *__restrict__ pred) {
  __shared__ uint32_t s_cnt[kThreads / 32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t first = blockIdx.x * blockDim.x, last = min(first + blockDim.x, n) - 1;
  const uint32_t s0 = sid[first];
  const bool uniform = s0 != kNoSeg && s0 == sid[last];
  const uint32_t s = i < n ? sid[i] : kNoSeg;
  bool p = false;
  if (s != kNoSeg && !w.leaf[s]) p = cs[(size_t)w.axis[s] * nt + idx[i]] < w.thresh[s];
  if (i < n) pred[i] = p ? 1 : 0;
  if (!uniform) {
    if (p) atomicAdd(&w.ntrue[s], 1u);
    return;
  }
  const unsigned b = __ballot_sync(0xFFFFFFFFu, p);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    for (int k = 0; k < kThreads / 32; k++) c += s_cnt[k];
    if (c) atomicAdd(&w.ntrue[s0], c);
  }
}

// misplaced flags, packed: high word = left part holds a "right" triangle, low word = right part holds a "left" one
// (one 64-bit scan then ranks both kinds)
__global__ void k_misplaced(uint32_t n, const uint32_t *__restrict__ sid, const uint32_t *__restrict__ sl,
                            const uint8_t *__restrict__ pred, SegWork w, unsigned long long *__restrict__ mis) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = sid[i];
  unsigned long long m = 0ull;
  if (s != kNoSeg && !w.leaf[s]) {
    const uint32_t mid = sl[s] + w.ntrue[s];
    if (i < mid) m = pred[i] ? 0ull : (1ull << 32);
    else m = pred[i] ? 1ull : 0ull;
  }
  mis[i] = m;
}

// ---- exclusive scan (three phases; out has n + 1 entries, out[n] = total) ------------------------------
constexpr int kScanItems = 8; // per thread
template <class V>
__global__ void __launch_bounds__(kThreads) k_scan_blocks(uint32_t n, const V *__restrict__ in, V *__restrict__ out,
                                                         V *__restrict__ sums) {
  __shared__ V s_w[kThreads / 32];
  const uint32_t base = blockIdx.x * (kThreads * kScanItems) + threadIdx.x * kScanItems;
  V v[kScanItems], t = 0;
  for (int k = 0; k < kScanItems; k++) {
    v[k] = (base + k < n) ? in[base + k] : (V)0;
    t += v[k];
  }
  V x = t; // inclusive scan of the per-thread totals
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const V y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if ((int)lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  V woff = 0;
  for (unsigned k = 0; k < warp; k++) woff += s_w[k];
  V run = woff + x - t;
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == kThreads - 1) sums[blockIdx.x] = woff + x;
}
template <class V>
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t nblocks, V *sums, V *total) { // <<<1, 1024>>>
  __shared__ V s_w[32];
  __shared__ V s_run;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const V v = i < nblocks ? sums[i] : (V)0;
    V x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const V y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    V woff = 0;
    for (unsigned k = 0; k < warp; k++) woff += s_w[k];
    const V run = s_run;
    if (i < nblocks) sums[i] = run + woff + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_run = run + woff + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_run;
}
template <class V>
__global__ void k_scan_add(uint32_t n, V *__restrict__ out, const V *__restrict__ sums, const V *__restrict__ total) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += sums[i / (kThreads * kScanItems)];
  if (i == 0) out[n] = *total;
}

// right-misplaced positions, compacted in position order
__global__ void k_right_list(uint32_t n, const unsigned long long *__restrict__ mis, const unsigned long long *__restrict__ scan,
                             uint32_t *__restrict__ rlist) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (uint32_t)mis[i]) rlist[(uint32_t)scan[i]] = i;
}

// k-th misplaced from the left <-> k-th misplaced from the right (libstdc++ bidirectional __partition)
__global__ void k_swap(uint32_t n, const uint32_t *__restrict__ sid, const uint32_t *__restrict__ sl,
                       const uint32_t *__restrict__ sr, const unsigned long long *__restrict__ mis,
                       const unsigned long long *__restrict__ scan, const uint32_t *__restrict__ rlist, uint32_t *idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !(mis[i] >> 32)) return;
  const uint32_t s = sid[i];
  const unsigned long long at_l = scan[sl[s]], at_r = scan[sr[s]];
  const uint32_t k = (uint32_t)(scan[i] >> 32) - (uint32_t)(at_l >> 32);
  const uint32_t m = (uint32_t)(at_r >> 32) - (uint32_t)(at_l >> 32); // misplaced on either side of this segment
  const uint32_t j = rlist[(uint32_t)at_l + (m - 1u - k)];
  const uint32_t a = idx[i], b = idx[j];
  idx[i] = b, idx[j] = a;
}

// ---- nodes of this level + segments of the next --------------------------------------------------------
struct NodeSoA {
  double *bmin, *bmax;   // [cap * 3]
  int32_t *flag, *axis;  // [cap]
  uint32_t *d0, *d1;     // branch: child node slots; leaf: ntris, first index
};

__global__ void k_emit_level(uint32_t S, Segs cur, SegWork w, const uint32_t *__restrict__ branch_ord, uint32_t node_base,
                             NodeSoA nodes, Segs next) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t slot = cur.node[s], l = cur.l[s], r = cur.r[s], n = r - l;
  for (int a = 0; a < 3; a++) nodes.bmin[3 * (size_t)slot + a] = w.bmin[3 * (size_t)s + a], nodes.bmax[3 * (size_t)slot + a] = w.bmax[3 * (size_t)s + a];
  if (w.leaf[s]) {
    nodes.flag[slot] = 1, nodes.axis[slot] = 0, nodes.d0[slot] = n, nodes.d1[slot] = l;
    return;
  }
  uint32_t mid = l + w.ntrue[s];
  if (mid == l || mid == r) mid = l + (n >> 1); // object-median fallback, array left as partitioned
  const uint32_t o = branch_ord[s];
  const uint32_t c0 = node_base + 2 * o, c1 = c0 + 1;
  nodes.flag[slot] = 0, nodes.axis[slot] = w.axis[s], nodes.d0[slot] = c0, nodes.d1[slot] = c1;
  next.l[2 * o] = l, next.r[2 * o] = mid, next.node[2 * o] = c0;
  next.l[2 * o + 1] = mid, next.r[2 * o + 1] = r, next.node[2 * o + 1] = c1;
}

// ---- pre-order numbering ---------------------------------------------------------------------------------
__global__ void k_sizes(uint32_t first, uint32_t count, NodeSoA nodes, uint32_t *size) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t s = first + i;
  size[s] = nodes.flag[s] ? 1u : 1u + size[nodes.d0[s]] + size[nodes.d1[s]];
}
__global__ void k_numbers(uint32_t first, uint32_t count, NodeSoA nodes, const uint32_t *size, uint32_t *num) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t s = first + i;
  if (nodes.flag[s]) return;
  num[nodes.d0[s]] = num[s] + 1u;
  num[nodes.d1[s]] = num[s] + 1u + size[nodes.d0[s]];
}
__global__ void k_emit_nodes(uint32_t count, NodeSoA nodes, const uint32_t *num, mb200_bvh_node *out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
  mb200_bvh_node nd;
  for (int a = 0; a < 3; a++) nd.bmin[a] = nodes.bmin[3 * (size_t)s + a], nd.bmax[a] = nodes.bmax[3 * (size_t)s + a];
  nd.flag = nodes.flag[s];
  nd.axis = nodes.axis[s];
  if (nd.flag) nd.data[0] = nodes.d0[s], nd.data[1] = nodes.d1[s];
  else nd.data[0] = num[nodes.d0[s]], nd.data[1] = num[nodes.d1[s]];
  out[num[s]] = nd;
}

// ---- device layout straight from the device tree (what scene.cc::relayout_bvh does on the host) --------
__global__ void k_branch_flags(uint32_t count, const mb200_bvh_node *__restrict__ nodes, uint32_t *__restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) flag[i] = nodes[i].flag == 0 ? 1u : 0u;
}
// pair_of[i] = number of branch nodes before node i in pre-order: relayout_bvh's numbering
__global__ void k_emit_pairs(uint32_t count, const mb200_bvh_node *__restrict__ nodes, const uint32_t *__restrict__ pair_of,
                             PairNode *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || nodes[i].flag != 0) return;
  PairNode p;
  for (int c = 0; c < 2; c++) {
    const uint32_t ci = nodes[i].data[c];
    const mb200_bvh_node ch = nodes[ci];
    for (int k = 0; k < 3; k++) p.box[c][k] = ch.bmin[k], p.box[c][3 + k] = ch.bmax[k];
    if (ch.flag == 0) p.ref[c] = pair_of[ci], p.cnt[c] = kBranch;
    else p.ref[c] = ch.data[1], p.cnt[c] = ch.data[0];
  }
  p.axis = (uint32_t)nodes[i].axis;
  p.pad_[0] = p.pad_[1] = p.pad_[2] = 0u;
  out[pair_of[i]] = p;
}
__global__ void k_emit_tris32(uint32_t nt, const uint32_t *__restrict__ idx, const double *__restrict__ v,
                              const uint32_t *__restrict__ f, const uint32_t *__restrict__ mat, TriRecordF32 *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const uint32_t t = idx[i];
  const double *a = v + 3 * (size_t)f[3 * (size_t)t + 0], *b = v + 3 * (size_t)f[3 * (size_t)t + 1], *c = v + 3 * (size_t)f[3 * (size_t)t + 2];
  TriRecordF32 r;
  for (int k = 0; k < 3; k++) r.p0[k] = (float)a[k], r.p1[k] = (float)b[k], r.p2[k] = (float)c[k];
  r.face = t;
  r.mat = mat ? mat[t] : 0xFFFFFFFFu; // bvh_accel.cc:685-689
  r.pad_ = 0u;
  out[i] = r;
}
__global__ void k_emit_tris64(uint32_t nt, const uint32_t *__restrict__ idx, const double *__restrict__ v,
                              const uint32_t *__restrict__ f, const uint32_t *__restrict__ mat, TriRecordF64 *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const uint32_t t = idx[i];
  const double *a = v + 3 * (size_t)f[3 * (size_t)t + 0], *b = v + 3 * (size_t)f[3 * (size_t)t + 1], *c = v + 3 * (size_t)f[3 * (size_t)t + 2];
  TriRecordF64 r;
  for (int k = 0; k < 3; k++) r.p0[k] = a[k], r.e1[k] = b[k] - a[k], r.e2[k] = c[k] - a[k]; // bvh_accel.cc:600-603
  r.face = t;
  r.mat = mat ? mat[t] : 0xFFFFFFFFu;
  out[i] = r;
}

inline unsigned grid_for(size_t n, int block = kThreads) { return (unsigned)((n + block - 1) / block); }

// One cudaMalloc carved into 256-byte aligned pieces.  The carving code runs twice: first to measure
// (base == nullptr), then for real.
struct Arena {
  char *base = nullptr;
  size_t used = 0;
  template <class T> void get(T **p, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    *p = base ? reinterpret_cast<T *>(base + used) : nullptr;
    used += bytes;
  }
  cudaError_t commit() {
    const size_t total = used;
    used = 0;
    return cudaMalloc((void **)&base, total ? total : 256);
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
  }
  ~Arena() { release(); }
};

template <class V> cudaError_t scan_excl(uint32_t n, const V *in, V *out, V *sums, V *total, cudaStream_t s) {
  const uint32_t per = kThreads * kScanItems, nblocks = (n + per - 1) / per;
  if (n == 0) return cudaMemsetAsync(out, 0, sizeof(V), s);
  k_scan_blocks<V><<<nblocks, kThreads, 0, s>>>(n, in, out, sums);
  k_scan_sums<V><<<1, 1024, 0, s>>>(nblocks, sums, total);
  k_scan_add<V><<<grid_for(n), kThreads, 0, s>>>(n, out, sums, total);
  return cudaGetLastError();
}

struct StageClock { // MB200_BUILD_TIMING=1: stage times on stderr (adds a synchronise per stage)
  bool on;
  cudaStream_t s;
  std::chrono::steady_clock::time_point prev;
  explicit StageClock(cudaStream_t st) : s(st) {
    const char *e = getenv("MB200_BUILD_TIMING");
    on = e && atoi(e) != 0;
    prev = std::chrono::steady_clock::now();
  }
  void operator()(const char *name) {
    if (!on) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mb200 build] %-10s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - prev).count());
    prev = now;
  }
};

// The tree on the device, reference layout: nodes in pre-order + the permuted index array, with the mesh arrays
// the build uploaded.  Lives until the arenas go.
struct DeviceTree {
  Arena work, result;
  double *d_v = nullptr;
  uint32_t *d_f = nullptr, *d_idx = nullptr;
  mb200_bvh_node *d_nodes = nullptr;
  uint32_t *d_scratch = nullptr, *d_scan = nullptr, *d_sums = nullptr, *d_total = nullptr; // node_count(+1) each
  uint32_t node_count = 0;
  mb200_build_stats stats{0, 0, 0};
  int launches = 0;
};

cudaError_t build_tree(DeviceTree &T, cudaStream_t s, const double *vertices, size_t nverts, const uint32_t *faces,
                       size_t nfaces, const mb200_build_options &opt, StageClock &stage) {
  const uint32_t nt = (uint32_t)nfaces;
  const int nb = opt.bin_size;
  // siblings together hold >= min_leaf_primitives triangles, so a level has at most 2 * nfaces / min_leaf segments
  const size_t max_segs = 2 * (nfaces / (size_t)opt.min_leaf_primitives) + 4;
  const size_t max_branch = nfaces / (size_t)opt.min_leaf_primitives + 2;
  const size_t max_nodes = 2 * nfaces + 2;
  double *d_lo, *d_hi, *d_cs;
  uint32_t *d_sid, *d_rlist, *d_sums, *d_total, *d_counter, *d_ord, *d_hist;
  unsigned long long *d_mis, *d_mscan, *d_sums64, *d_total64;
  uint8_t *d_pred;
  Segs seg[2];
  SegWork w;
  NodeSoA nodes;
  auto carve = [&](Arena &A) {
    A.get(&T.d_v, 3 * nverts);
    A.get(&T.d_f, 3 * nfaces);
    A.get(&T.d_idx, nfaces);
    A.get(&d_lo, 3 * nfaces);
    A.get(&d_hi, 3 * nfaces);
    A.get(&d_cs, 3 * nfaces);
    A.get(&d_sid, nfaces);
    A.get(&d_mis, nfaces);
    A.get(&d_mscan, nfaces + 1);
    A.get(&d_sums64, nfaces / (kThreads * kScanItems) + 2);
    A.get(&d_total64, 1);
    A.get(&d_rlist, nfaces);
    A.get(&d_pred, nfaces);
    A.get(&d_sums, max_segs / (kThreads * kScanItems) + 2);
    A.get(&d_total, 1);
    A.get(&d_counter, 1);
    A.get(&d_ord, max_segs + 1);
    for (int k = 0; k < 2; k++) A.get(&seg[k].l, max_segs), A.get(&seg[k].r, max_segs), A.get(&seg[k].node, max_segs);
    A.get(&w.kmin, 3 * max_segs);
    A.get(&w.kmax, 3 * max_segs);
    A.get(&w.bmin, 3 * max_segs);
    A.get(&w.bmax, 3 * max_segs);
    A.get(&w.scale, 3 * max_segs);
    A.get(&w.step, 3 * max_segs);
    A.get(&w.leaf, max_segs);
    A.get(&w.hslot, max_segs);
    A.get(&w.axis, max_segs);
    A.get(&w.thresh, max_segs);
    A.get(&w.ntrue, max_segs);
    A.get(&w.isbranch, max_segs);
    A.get(&nodes.bmin, 3 * max_nodes);
    A.get(&nodes.bmax, 3 * max_nodes);
    A.get(&nodes.flag, max_nodes);
    A.get(&nodes.axis, max_nodes);
    A.get(&nodes.d0, max_nodes);
    A.get(&nodes.d1, max_nodes);
    A.get(&d_hist, max_branch * 6 * (size_t)nb);
  };
  carve(T.work);
  CUB(T.work.commit());
  carve(T.work);
  stage("alloc");

  CUB(cudaMemcpyAsync(T.d_v, vertices, 3 * nverts * sizeof(double), cudaMemcpyHostToDevice, s));
  CUB(cudaMemcpyAsync(T.d_f, faces, 3 * nfaces * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  k_tri_cache<<<grid_for(nt), kThreads, 0, s>>>(T.d_v, T.d_f, nt, d_lo, d_hi, d_cs, T.d_idx);
  int nl = 1;
  stage("upload");

  // level 0: one segment, node slot 0
  const uint32_t zero = 0;
  CUB(cudaMemcpyAsync(seg[0].l, &zero, 4, cudaMemcpyHostToDevice, s));
  CUB(cudaMemcpyAsync(seg[0].r, &nt, 4, cudaMemcpyHostToDevice, s));
  CUB(cudaMemcpyAsync(seg[0].node, &zero, 4, cudaMemcpyHostToDevice, s));
  std::vector<uint32_t> level_first, level_count; // node slots of every level
  uint32_t S = 1, node_count = 1;
  int level = 0, cur = 0;
  size_t leaves = 0, branches = 0;
  while (S > 0) {
    if ((size_t)S > max_segs) return cudaErrorInvalidValue;
    level_first.push_back(node_count - S);
    level_count.push_back(S);
    const Segs c = seg[cur], nx = seg[cur ^ 1];
    k_assign_sid<<<grid_for(nt), kThreads, 0, s>>>(nt, c.l, c.r, S, d_sid);
    k_init_keys<<<grid_for(3 * (size_t)S), kThreads, 0, s>>>(3 * S, w.kmin, w.kmax);
    k_bounds<<<grid_for(nt), kThreads, 0, s>>>(nt, nt, d_sid, T.d_idx, d_lo, d_hi, w.kmin, w.kmax);
    CUB(cudaMemsetAsync(d_counter, 0, 4, s));
    k_seg_setup<<<grid_for(S), kThreads, 0, s>>>(S, c.l, c.r, level, opt, w, d_counter);
    uint32_t nbranch = 0;
    CUB(cudaMemcpyAsync(&nbranch, d_counter, 4, cudaMemcpyDeviceToHost, s));
    CUB(cudaStreamSynchronize(s));
    nl += 5;
    if (nbranch > 0) {
      if ((size_t)nbranch > max_branch) return cudaErrorInvalidValue;
      CUB(cudaMemsetAsync(d_hist, 0, (size_t)nbranch * 6 * nb * sizeof(uint32_t), s));
      const size_t hsm = (nb <= 512) ? 6 * (size_t)nb * sizeof(uint32_t) : 0;
      k_histogram<<<grid_for(nt), kThreads, hsm, s>>>(nt, nt, nb, d_sid, T.d_idx, d_lo, d_hi, w, d_hist);
      k_sweep<<<grid_for((size_t)S * 32), kThreads, 0, s>>>(S, c.l, c.r, opt, w, d_hist);
      k_pred<<<grid_for(nt), kThreads, 0, s>>>(nt, nt, d_sid, T.d_idx, d_cs, w, d_pred);
      k_misplaced<<<grid_for(nt), kThreads, 0, s>>>(nt, d_sid, c.l, d_pred, w, d_mis);
      CUB(scan_excl(nt, d_mis, d_mscan, d_sums64, d_total64, s));
      k_right_list<<<grid_for(nt), kThreads, 0, s>>>(nt, d_mis, d_mscan, d_rlist);
      k_swap<<<grid_for(nt), kThreads, 0, s>>>(nt, d_sid, c.l, c.r, d_mis, d_mscan, d_rlist, T.d_idx);
      nl += 9;
    }
    CUB(scan_excl(S, w.isbranch, d_ord, d_sums, d_total, s));
    k_emit_level<<<grid_for(S), kThreads, 0, s>>>(S, c, w, d_ord, node_count, nodes, nx);
    nl += 4;
    CUB(cudaGetLastError());
    leaves += S - nbranch;
    branches += nbranch;
    node_count += 2 * nbranch;
    S = 2 * nbranch;
    cur ^= 1;
    level++;
    if (level > opt.max_tree_depth + 1) return cudaErrorInvalidValue; // cannot happen: depth >= max_tree_depth makes leaves
  }
  stage("levels");

  // pre-order numbers: sizes bottom-up, numbers top-down (left child = parent + 1), then emit
  uint32_t *d_size, *d_num;
  auto carve2 = [&](Arena &A) {
    A.get(&T.d_nodes, node_count);
    A.get(&d_size, node_count);
    A.get(&d_num, node_count);
    A.get(&T.d_scan, (size_t)node_count + 1);
    A.get(&T.d_sums, node_count / (kThreads * kScanItems) + 2);
    A.get(&T.d_total, 1);
  };
  carve2(T.result);
  CUB(T.result.commit());
  carve2(T.result);
  T.d_scratch = d_size; // free again once the numbers are out
  for (int lv = (int)level_first.size() - 1; lv >= 0; lv--) {
    k_sizes<<<grid_for(level_count[lv]), kThreads, 0, s>>>(level_first[lv], level_count[lv], nodes, d_size);
    nl++;
  }
  CUB(cudaMemsetAsync(d_num, 0, sizeof(uint32_t), s));
  for (size_t lv = 0; lv < level_first.size(); lv++) {
    k_numbers<<<grid_for(level_count[lv]), kThreads, 0, s>>>(level_first[lv], level_count[lv], nodes, d_size, d_num);
    nl++;
  }
  k_emit_nodes<<<grid_for(node_count), kThreads, 0, s>>>(node_count, nodes, d_num, T.d_nodes);
  nl++;
  CUB(cudaGetLastError());
  stage("numbering");
  T.node_count = node_count;
  T.stats.max_tree_depth = (int)level_first.size() - 1;
  T.stats.num_leaf_nodes = (int)leaves;
  T.stats.num_branch_nodes = (int)branches;
  T.launches = nl;
  return cudaSuccess;
}

cudaError_t download_tree(const DeviceTree &T, HostBVH &out, size_t nfaces, cudaStream_t s) {
  out.nodes.resize(T.node_count);
  out.indices.resize(nfaces);
  CUB(cudaMemcpyAsync(out.nodes.data(), T.d_nodes, (size_t)T.node_count * sizeof(mb200_bvh_node), cudaMemcpyDeviceToHost, s));
  CUB(cudaMemcpyAsync(out.indices.data(), T.d_idx, nfaces * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUB(cudaStreamSynchronize(s));
  out.stats = T.stats;
  return cudaSuccess;
}

bool check_build_args(const uint32_t *faces, size_t nverts, size_t nfaces, const mb200_build_options &opt, std::string *err) {
  if (opt.bin_size <= 1 || opt.bin_size > 65536) {
    if (err) *err = "bin_size must be in (1, 65536]";
    return false;
  }
  if (opt.min_leaf_primitives < 2) {
    if (err) *err = "the device builder needs min_leaf_primitives >= 2 (use mb200_bvh_build)";
    return false;
  }
  if (nfaces > 0x7FFFFFF0ull) {
    if (err) *err = "too many triangles for the device builder";
    return false;
  }
  // one histogram per open branch node of a level: (nfaces / min_leaf) x 3 axes x 2 x bin_size counters
  if ((double)(nfaces / (size_t)opt.min_leaf_primitives + 2) * 6.0 * opt.bin_size * sizeof(uint32_t) > 16e9) {
    if (err) *err = "bin_size x triangle count needs more than 16 GB of histograms on the device (use mb200_bvh_build)";
    return false;
  }
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (long i = 0; i < (long)(3 * nfaces); i++) ok = ok && faces[i] < nverts;
  if (!ok && err) *err = "face references a vertex out of range";
  return ok;
}

} // namespace

// Entry used by the C ABI (mb200_bvh_build_device).  Validation as host/bvh_build.cc.
bool build_bvh_device(HostBVH &out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                      size_t nfaces, const mb200_build_options &opt, std::string *err, bool *cuda_failure) {
  if (cuda_failure) *cuda_failure = false;
  out.nodes.clear();
  out.indices.clear();
  out.stats = mb200_build_stats{0, 0, 0};
  if (!check_build_args(faces, nverts, nfaces, opt, err)) return false;
  if (nfaces == 0) return true; // empty tree: every ray misses
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) {
    cudaStream_t s = nullptr; // default stream: a one-off set-up step
    StageClock stage(s);
    DeviceTree T;
    e = build_tree(T, s, vertices, nverts, faces, nfaces, opt, stage);
    if (e == cudaSuccess) e = download_tree(T, out, nfaces, s);
    stage("download");
    T.work.release();
    T.result.release();
    stage("free");
  }
  if (e != cudaSuccess) {
    if (err) *err = std::string("device BVH build: ") + cudaGetErrorString(e);
    if (cuda_failure) *cuda_failure = true;
    out.nodes.clear();
    out.indices.clear();
    cudaGetLastError();
    return false;
  }
  return true;
}

// BVHAccel::Build + the scene upload in one step (mb200_scene_build): the tree is grown on the device and the
// traversal layout (layout.h) is written from it there; only the mesh goes up and, if asked for, the
// reference-layout tree comes down.  The result equals scene_create(build_bvh(...)) byte for byte.
int scene_build_device(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                       size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                       const mb200_build_options &opt, HostBVH *bvh_out, std::string *err) {
  *out = nullptr;
  if (bvh_out) {
    bvh_out->nodes.clear();
    bvh_out->indices.clear();
    bvh_out->stats = mb200_build_stats{0, 0, 0};
  }
  if (!check_build_args(faces, nverts, nfaces, opt, err)) return MB200_ERR_INVALID_ARG;
  if (nfaces == 0) // empty scene: every ray misses
    return scene_create(out, device, vertices, nverts, faces, 0, material_ids, fv_normals, fv_uvs, nullptr, 0, nullptr, 0, err);
  mb200_scene *sc = nullptr;
  int st = scene_open(&sc, device, err);
  if (st != MB200_OK) return st;
  cudaStream_t s = sc->stream;
  StageClock stage(s);
  cudaError_t e = cudaSuccess;
  auto cuda_fail = [&](const char *what) {
    if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    scene_destroy(sc);
    return (int)(e == cudaErrorMemoryAllocation ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA);
  };
  DeviceTree T;
  if ((e = build_tree(T, s, vertices, nverts, faces, nfaces, opt, stage)) != cudaSuccess) return cuda_fail("device BVH build");

  if (T.stats.max_tree_depth > 500) { // as relayout_bvh: the traversal stack holds 512 entries
    if (err) *err = "BVH deeper than 500 levels (reference stack is 512, bvh_accel.cc:548)";
    scene_destroy(sc);
    return MB200_ERR_INVALID_ARG;
  }
  const bool f32 = choose_tri_f32(vertices, 3 * nverts);
  const uint32_t nt = (uint32_t)nfaces, npairs = (uint32_t)T.stats.num_branch_nodes;
  auto dev_alloc = [&](void **p, size_t bytes) {
    e = cudaMalloc(p, bytes ? bytes : 16);
    if (e == cudaSuccess) sc->allocs.push_back(*p), sc->device_bytes += bytes;
    return e == cudaSuccess;
  };
  PairNode *d_pairs = nullptr;
  void *d_tris = nullptr, *d_v = nullptr, *d_f = nullptr, *d_n = nullptr, *d_uv = nullptr, *d_mat = nullptr;
  if (!dev_alloc((void **)&d_pairs, (size_t)npairs * sizeof(PairNode)) ||
      !dev_alloc(&d_tris, nfaces * (f32 ? sizeof(TriRecordF32) : sizeof(TriRecordF64))) ||
      !dev_alloc(&d_v, 3 * nverts * sizeof(double)) || !dev_alloc(&d_f, 3 * nfaces * sizeof(uint32_t)) ||
      (fv_normals && !dev_alloc(&d_n, 9 * nfaces * sizeof(double))) || (fv_uvs && !dev_alloc(&d_uv, 6 * nfaces * sizeof(double))))
    return cuda_fail("cudaMalloc");
  if (material_ids) {
    if ((e = cudaMalloc(&d_mat, nfaces * sizeof(uint32_t))) != cudaSuccess) return cuda_fail("cudaMalloc");
    e = cudaMemcpyAsync(d_mat, material_ids, nfaces * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, T.d_v, 3 * nverts * sizeof(double), cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, T.d_f, 3 * nfaces * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess && d_n) e = cudaMemcpyAsync(d_n, fv_normals, 9 * nfaces * sizeof(double), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && d_uv) e = cudaMemcpyAsync(d_uv, fv_uvs, 6 * nfaces * sizeof(double), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    k_branch_flags<<<grid_for(T.node_count), kThreads, 0, s>>>(T.node_count, T.d_nodes, T.d_scratch);
    e = scan_excl(T.node_count, T.d_scratch, T.d_scan, T.d_sums, T.d_total, s);
  }
  if (e == cudaSuccess) {
    k_emit_pairs<<<grid_for(T.node_count), kThreads, 0, s>>>(T.node_count, T.d_nodes, T.d_scan, d_pairs);
    if (f32)
      k_emit_tris32<<<grid_for(nt), kThreads, 0, s>>>(nt, T.d_idx, T.d_v, T.d_f, (const uint32_t *)d_mat, (TriRecordF32 *)d_tris);
    else
      k_emit_tris64<<<grid_for(nt), kThreads, 0, s>>>(nt, T.d_idx, T.d_v, T.d_f, (const uint32_t *)d_mat, (TriRecordF64 *)d_tris);
    e = cudaGetLastError();
  }
  mb200_bvh_node root;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&root, T.d_nodes, sizeof(root), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  stage("layout");
  if (e == cudaSuccess && bvh_out) e = download_tree(T, *bvh_out, nfaces, s);
  stage("download");
  if (d_mat) cudaFree(d_mat);
  T.work.release();
  T.result.release();
  stage("free");
  if (e != cudaSuccess) return cuda_fail("device scene build");

  SceneView &v = sc->view;
  v.empty = 0;
  v.tri_f32 = f32 ? 1 : 0;
  v.num_vertices = (uint32_t)nverts;
  v.num_faces = nt;
  v.nodes = d_pairs;
  v.tris = d_tris;
  v.num_pair_nodes = npairs;
  v.num_tris = nt;
  if (root.flag == 0) v.root_ref = 0u, v.root_cnt = kBranch;
  else v.root_ref = root.data[1], v.root_cnt = root.data[0];
  for (int k = 0; k < 3; k++) {
    v.root_box[k] = sc->root_bmin[k] = root.bmin[k];
    v.root_box[3 + k] = sc->root_bmax[k] = root.bmax[k];
  }
  v.vertices = (const double *)d_v;
  v.faces = (const uint32_t *)d_f;
  v.fv_normals = (const double *)d_n;
  v.fv_uvs = (const double *)d_uv;
  sc->tree_depth = T.stats.max_tree_depth;
  sc->stack_cap = T.stats.max_tree_depth + 2;
  st = scene_finish(sc, err);
  if (st != MB200_OK) {
    scene_destroy(sc);
    return st;
  }
  *out = sc;
  return MB200_OK;
}

} // namespace mb200

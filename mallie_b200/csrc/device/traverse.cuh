// traverse.cuh -- device-side primitives of BVHAccel::Traverse (bvh_accel.cc:546-844) for sm_100a:
// ray set-up, the slab test, the triangle test, triangle-record loaders and the traversal stack.
// Compiled with -fmad=false: every multiply and add below rounds separately, in the reference's
// evaluation order, so hit records are bit-identical to the CPU reference.  The traversal loop
// itself is the state machine in trace_sm.cuh.
//
// Equivalence of the PairNode walk with the reference loop (bvh_accel.cc:805-834), which pops a
// node, box-tests it against the *current* hitT, and then pushes far/near children:
//   * here both child boxes are tested when their parent is visited (they sit in the parent's
//     128-byte PairNode).  The three slab conditions
//         (tmax > 0) && (tmin <= tmax) && (tmin <= hitT)
//     are evaluated with the same arithmetic; only the last one depends on hitT.
//   * hitT never increases, so a child that fails (tmin <= hitT) now also fails it at the
//     reference's later pop time: dropping it immediately is exact.
//   * a child that passes is either continued with immediately (near child, hitT unchanged => the
//     reference's pop-time test passes too) or pushed together with its tmin; when it is popped
//     the remaining condition (tmin <= hitT_now) is re-checked -- the reference's pop-time decision.
//   * near/far order is sign[axis] as in bvh_accel.cc:818-823; triangles are stored in indices_
//     order, so "last accepted among equal t wins" is kept.
// Consequently the sequence of leaves visited, triangles tested and the number of box tests
// (1 + 2 per accepted branch) are those of the reference.
#ifndef MALLIE_B200_TRAVERSE_CUH_
#define MALLIE_B200_TRAVERSE_CUH_

#include <cfloat>
#include <cuda_runtime.h>

#include "layout.h"

namespace mb200 {

// std::numeric_limits<double>::epsilon() * 1024 (bvh_accel.cc:598)
#define MB200_TRI_EPS (2.2204460492503131e-16 * 1024.0)

struct RayD {
  double ox, oy, oz;
  double dx, dy, dz;
  double ix, iy, iz; // 1.0 / dir, may be +-inf (bvh_accel.cc:793-797)
  uint32_t sgn;      // bit a = (dir[a] < 0.0), so -0.0 counts as non-negative (bvh_accel.cc:787-790); one
                     // register instead of three flags (which ptxas spilled under the 64-register cap)
};

__device__ __forceinline__ void ray_setup(RayD &r, double ox, double oy, double oz, double dx, double dy, double dz) {
  r.ox = ox, r.oy = oy, r.oz = oz;
  r.dx = dx, r.dy = dy, r.dz = dz;
  r.sgn = (dx < 0.0 ? 1u : 0u) | (dy < 0.0 ? 2u : 0u) | (dz < 0.0 ? 4u : 0u);
  r.ix = 1.0 / dx, r.iy = 1.0 / dy, r.iz = 1.0 / dz;
}

// IntersectRayAABB (bvh_accel.cc:550-593).  Ternaries, not fmin/fmax: NaNs from
// 0*inf propagate exactly as in the reference.  b = {bmin[3], bmax[3]}.
__device__ __forceinline__ bool slab_test(const double b0, const double b1, const double b2, const double b3,
                                          const double b4, const double b5, const RayD &r, const double max_t,
                                          double &tmin_out) {
  const bool sx = (r.sgn & 1u) != 0u, sy = (r.sgn & 2u) != 0u, sz = (r.sgn & 4u) != 0u;
  const double min_x = sx ? b3 : b0;
  const double min_y = sy ? b4 : b1;
  const double min_z = sz ? b5 : b2;
  const double max_x = sx ? b0 : b3;
  const double max_y = sy ? b1 : b4;
  const double max_z = sz ? b2 : b5;

  const double tmin_x = (min_x - r.ox) * r.ix;
  const double tmax_x = (max_x - r.ox) * r.ix;
  const double tmin_y = (min_y - r.oy) * r.iy;
  const double tmax_y = (max_y - r.oy) * r.iy;

  double tmin = (tmin_x > tmin_y) ? tmin_x : tmin_y;
  double tmax = (tmax_x < tmax_y) ? tmax_x : tmax_y;

  const double tmin_z = (min_z - r.oz) * r.iz;
  const double tmax_z = (max_z - r.oz) * r.iz;

  tmin = (tmin > tmin_z) ? tmin : tmin_z;
  tmax = (tmax < tmax_z) ? tmax : tmax_z;

  tmin_out = tmin;
  return (tmax > 0.0) && (tmin <= tmax) && (tmin <= max_t);
}

struct HitD {
  double t, u, v;
  uint32_t face, mat;
};

// ---- triangle records (layout.h) ----------------------------------------------------------------
// Both record kinds are presented to the triangle test as p0 plus the two edges e1 = p1 - p0,
// e2 = p2 - p0.  The f64 record stores the edges, rounded once on the host exactly as TriangleIsect
// rounds them (one IEEE double subtraction each, bvh_accel.cc:600-603); the f32 record stores the
// float-exact vertices and the edges are formed here after widening to double (same result).
struct TriEdges {
  double p0x, p0y, p0z, e1x, e1y, e1z, e2x, e2y, e2z;
  uint32_t face, mat;
};

template <bool F32> __device__ __forceinline__ TriEdges load_tri_edges(const void *tris, uint32_t i);

template <> __device__ __forceinline__ TriEdges load_tri_edges<true>(const void *tris, uint32_t i) {
  const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const TriRecordF32 *>(tris) + i);
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  TriEdges t;
  t.p0x = (double)a.x, t.p0y = (double)a.y, t.p0z = (double)a.z;
  t.e1x = (double)b.x - t.p0x, t.e1y = (double)b.y - t.p0y, t.e1z = (double)b.z - t.p0z;
  t.e2x = (double)c.x - t.p0x, t.e2y = (double)c.y - t.p0y, t.e2z = (double)c.z - t.p0z;
  t.face = __float_as_uint(a.w);
  t.mat = __float_as_uint(b.w);
  return t;
}

template <> __device__ __forceinline__ TriEdges load_tri_edges<false>(const void *tris, uint32_t i) {
  const double2 *p = reinterpret_cast<const double2 *>(reinterpret_cast<const TriRecordF64 *>(tris) + i);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
  TriEdges t;
  t.p0x = a.x, t.p0y = a.y, t.p0z = b.x;
  t.e1x = b.y, t.e1y = c.x, t.e1z = c.y;
  t.e2x = d.x, t.e2y = d.y, t.e2z = e.x;
  const unsigned long long w = (unsigned long long)__double_as_longlong(e.y);
  t.face = (uint32_t)(w & 0xFFFFFFFFull);
  t.mat = (uint32_t)(w >> 32);
  return t;
}

// TriangleIsect (bvh_accel.cc:595-638): Moeller-Trumbore, no culling.
// BRANCH_FREE: the same arithmetic with every rejection test folded into one predicate.  Each test is
// written as the negation of the reference's rejecting comparison, so NaNs fall through exactly as they
// do there (a NaN u, v or t is NOT rejected by `u < 0.0 || u > 1.0` and friends); a near-zero det still
// rejects first, whatever 1.0 / det produced.
template <bool BRANCH_FREE>
__device__ __forceinline__ bool tri_test_edges(double &t_io, double &u_out, double &v_out, const TriEdges &k,
                                               const RayD &r) {
  // p = dir x e2
  const double px = r.dy * k.e2z - r.dz * k.e2y;
  const double py = r.dz * k.e2x - r.dx * k.e2z;
  const double pz = r.dx * k.e2y - r.dy * k.e2x;
  const double det = k.e1x * px + k.e1y * py + k.e1z * pz;
  if (!BRANCH_FREE && fabs(det) < MB200_TRI_EPS) return false;
  const double inv_det = 1.0 / det;
  const double sx = r.ox - k.p0x, sy = r.oy - k.p0y, sz = r.oz - k.p0z;
  // q = s x e1
  const double qx = sy * k.e1z - sz * k.e1y;
  const double qy = sz * k.e1x - sx * k.e1z;
  const double qz = sx * k.e1y - sy * k.e1x;
  const double u = (sx * px + sy * py + sz * pz) * inv_det;
  const double v = (qx * r.dx + qy * r.dy + qz * r.dz) * inv_det;
  const double t = (k.e2x * qx + k.e2y * qy + k.e2z * qz) * inv_det;
  if (BRANCH_FREE) {
    const bool ok = !(fabs(det) < MB200_TRI_EPS) & !(u < 0.0) & !(u > 1.0) & !(v < 0.0) & !(u + v > 1.0) & !(t < 0.0) &
                    !(t > t_io);
    if (!ok) return false;
  } else {
    if (u < 0.0 || u > 1.0) return false;
    if (v < 0.0 || u + v > 1.0) return false;
    if (t < 0.0 || t > t_io) return false;
  }
  t_io = t;
  u_out = u;
  v_out = v;
  return true;
}

// ---- traversal stack: first S entries per thread in shared memory (column layout: entry k of
// thread t at [k * blockDim.x + t] -> conflict-free 128-bit accesses), the rest in a per-thread
// local-memory array.  An entry is (tmin of the pushed child, its ref, its cnt).
template <int S, int CAP> struct TravStack {
  uint32_t sm_addr; // this thread's column base as a shared-window address (kept in one register: without
                    // it ptxas re-derived the pointer from %tid and the CTA id on every pop)
  uint32_t stride_bytes;
  uint4 ovf[(CAP > S) ? (CAP - S) : 1];
  __device__ __forceinline__ void init(uint4 *column, int stride) {
    sm_addr = (uint32_t)__cvta_generic_to_shared(column);
    stride_bytes = (uint32_t)stride * 16u;
  }
  __device__ __forceinline__ void put(int k, double tmin, uint32_t ref, uint32_t cnt) {
    const unsigned long long tb = (unsigned long long)__double_as_longlong(tmin);
    const uint4 e = make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), ref, cnt);
    if (S > 0 && k < S)
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sm_addr + (uint32_t)k * stride_bytes), "r"(e.x), "r"(e.y),
                   "r"(e.z), "r"(e.w)
                   : "memory");
    else ovf[k - S] = e;
  }
  __device__ __forceinline__ void get(int k, double &tmin, uint32_t &ref, uint32_t &cnt) const {
    uint4 e;
    if (S > 0 && k < S)
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                   : "r"(sm_addr + (uint32_t)k * stride_bytes)
                   : "memory");
    else e = ovf[k - S];
    tmin = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
    ref = e.z;
    cnt = e.w;
  }
};

struct TravCounters {
  unsigned int nodes, tris, max_stack;
};

} // namespace mb200

#endif

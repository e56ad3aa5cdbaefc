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

// The same test with the ray's direction signs known at compile time (SGN bit a = dir[a] < 0): the six
// per-axis selects of slab_test disappear.  Used by the inner step when all of its lanes share an octant
// (trace_sm.cuh, kVarOctant); the arithmetic and the comparisons are slab_test's, value for value.
template <int SGN>
__device__ __forceinline__ bool slab_test_oct(const double (&b)[6], const RayD &r, const double max_t, double &tmin_out) {
  const double min_x = (SGN & 1) ? b[3] : b[0], max_x = (SGN & 1) ? b[0] : b[3];
  const double min_y = (SGN & 2) ? b[4] : b[1], max_y = (SGN & 2) ? b[1] : b[4];
  const double min_z = (SGN & 4) ? b[5] : b[2], max_z = (SGN & 4) ? b[2] : b[5];
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

// ---- triangle records (layout.h) ----------------------------------------------------------------
// Every record kind is presented to the triangle test as p0 plus the two edges e1 = p1 - p0,
// e2 = p2 - p0.  The f64 records store the edges, rounded once on the host / by the layout kernel exactly as
// TriangleIsect rounds them (one IEEE double subtraction each, bvh_accel.cc:600-603); the f32 records store
// the float-exact vertices and the edges are formed here after widening to double (same result).
struct TriEdges {
  double p0x, p0y, p0z, e1x, e1y, e1z, e2x, e2y, e2z;
  uint32_t face, mat;
};

template <int KIND> __device__ __forceinline__ TriEdges load_tri_edges(const void *tris, uint32_t i);

__device__ __forceinline__ TriEdges tri_edges_from_f32(const float4 a, const float4 b, const float4 c) {
  TriEdges t;
  t.p0x = (double)a.x, t.p0y = (double)a.y, t.p0z = (double)a.z;
  t.e1x = (double)b.x - t.p0x, t.e1y = (double)b.y - t.p0y, t.e1z = (double)b.z - t.p0z;
  t.e2x = (double)c.x - t.p0x, t.e2y = (double)c.y - t.p0y, t.e2z = (double)c.z - t.p0z;
  t.face = __float_as_uint(a.w);
  t.mat = __float_as_uint(b.w);
  return t;
}

template <> __device__ __forceinline__ TriEdges load_tri_edges<kTriF32>(const void *tris, uint32_t i) {
  const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const TriRecordF32 *>(tris) + i);
  return tri_edges_from_f32(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

template <> __device__ __forceinline__ TriEdges load_tri_edges<kTriF64>(const void *tris, uint32_t i) {
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

// 64-byte f32 record: two 256-bit loads (two L1 wavefronts per lane instead of three; c4..c7 are the record's padding)
#pragma nv_diag_suppress 550
template <> __device__ __forceinline__ TriEdges load_tri_edges<kTriF32x64>(const void *tris, uint32_t i) {
  const char *p = reinterpret_cast<const char *>(tris) + (size_t)i * 64u;
  float a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3, c4, c5, c6, c7;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
               : "l"(p));
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(c0), "=f"(c1), "=f"(c2), "=f"(c3), "=f"(c4), "=f"(c5), "=f"(c6), "=f"(c7)
               : "l"(p + 32));
  return tri_edges_from_f32(make_float4(a0, a1, a2, a3), make_float4(b0, b1, b2, b3), make_float4(c0, c1, c2, c3));
}

#pragma nv_diag_default 550

// 96-byte f64 record: three 256-bit loads, no conversions and no edge subtractions
template <> __device__ __forceinline__ TriEdges load_tri_edges<kTriF64x96>(const void *tris, uint32_t i) {
  const char *p = reinterpret_cast<const char *>(tris) + (size_t)i * 96u;
  double v[12];
#pragma unroll
  for (int k = 0; k < 3; k++)
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[4 * k]), "=d"(v[4 * k + 1]), "=d"(v[4 * k + 2]), "=d"(v[4 * k + 3])
                 : "l"(p + 32 * k));
  TriEdges t;
  t.p0x = v[0], t.p0y = v[1], t.p0z = v[2];
  t.e1x = v[3], t.e1y = v[4], t.e1z = v[5];
  t.e2x = v[6], t.e2y = v[7], t.e2z = v[8];
  const unsigned long long w = (unsigned long long)__double_as_longlong(v[9]);
  t.face = (uint32_t)(w & 0xFFFFFFFFull);
  t.mat = (uint32_t)(w >> 32);
  return t;
}

// ---- Woop's test (development variant, NOT bit-exact with TriangleIsect) -------------------------------------
// The record maps world space onto the triangle's own frame, in which it is (0,0,0) (1,0,0) (0,1,0):
//   r3 = n = e1 x e2 (so r3 . d = -det of TriangleIsect and the same |det| < eps rejection applies),
//   r1 = (e2 x n) / (e1 . (e2 x n)),  r2 = (n x e1) / (e2 . (n x e1)),  b_k = -r_k . p0.
//   t = -(r3 . o + b3) / (r3 . d),  u = (r1 . o + b1) + t (r1 . d),  v = (r2 . o + b2) + t (r2 . d)
// 20 multiplies, 17 additions and one division against Moeller-Trumbore's 27 + 17 (edges precomputed) + one.
struct WoopRec {
  double r1x, r1y, r1z, b1, r2x, r2y, r2z, b2, r3x, r3y, r3z, b3;
};
__device__ __forceinline__ WoopRec load_woop(const void *tris, uint32_t i) {
  const char *p = reinterpret_cast<const char *>(tris) + (size_t)i * 96u;
  double v[12];
#pragma unroll
  for (int k = 0; k < 3; k++)
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[4 * k]), "=d"(v[4 * k + 1]), "=d"(v[4 * k + 2]), "=d"(v[4 * k + 3])
                 : "l"(p + 32 * k));
  return WoopRec{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]};
}
__device__ __forceinline__ bool tri_test_woop(double &t_io, double &u_out, double &v_out, const WoopRec &w, const RayD &r) {
  const double dz = (w.r3x * r.dx + w.r3y * r.dy) + w.r3z * r.dz;
  if (fabs(dz) < MB200_TRI_EPS) return false;
  const double oz = ((w.r3x * r.ox + w.r3y * r.oy) + w.r3z * r.oz) + w.b3;
  const double t = -oz / dz;
  const double ox = ((w.r1x * r.ox + w.r1y * r.oy) + w.r1z * r.oz) + w.b1;
  const double dx = (w.r1x * r.dx + w.r1y * r.dy) + w.r1z * r.dz;
  const double u = ox + t * dx;
  const double oy = ((w.r2x * r.ox + w.r2y * r.oy) + w.r2z * r.oz) + w.b2;
  const double dy = (w.r2x * r.dx + w.r2y * r.dy) + w.r2z * r.dz;
  const double v = oy + t * dy;
  if (u < 0.0 || u > 1.0) return false;
  if (v < 0.0 || u + v > 1.0) return false;
  if (t < 0.0 || t > t_io) return false;
  t_io = t;
  u_out = u;
  v_out = v;
  return true;
}
// faceID / materialID of canonical record i (read on accept only)
__device__ __forceinline__ void tri_ids(const void *tris, int tri_f32, uint32_t i, uint32_t &face, uint32_t &mat) {
  if (tri_f32) {
    const TriRecordF32 *t = reinterpret_cast<const TriRecordF32 *>(tris) + i;
    face = __ldg(&t->face), mat = __ldg(&t->mat);
  } else {
    const TriRecordF64 *t = reinterpret_cast<const TriRecordF64 *>(tris) + i;
    face = __ldg(&t->face), mat = __ldg(&t->mat);
  }
}

// TriangleIsect (bvh_accel.cc:595-638): Moeller-Trumbore, no culling.
__device__ __forceinline__ bool tri_test_edges(double &t_io, double &u_out, double &v_out, const TriEdges &k,
                                               const RayD &r) {
  // p = dir x e2
  const double px = r.dy * k.e2z - r.dz * k.e2y;
  const double py = r.dz * k.e2x - r.dx * k.e2z;
  const double pz = r.dx * k.e2y - r.dy * k.e2x;
  const double det = k.e1x * px + k.e1y * py + k.e1z * pz;
  if (fabs(det) < MB200_TRI_EPS) return false;
  const double inv_det = 1.0 / det;
  const double sx = r.ox - k.p0x, sy = r.oy - k.p0y, sz = r.oz - k.p0z;
  // q = s x e1
  const double qx = sy * k.e1z - sz * k.e1y;
  const double qy = sz * k.e1x - sx * k.e1z;
  const double qz = sx * k.e1y - sy * k.e1x;
  const double u = (sx * px + sy * py + sz * pz) * inv_det;
  const double v = (qx * r.dx + qy * r.dy + qz * r.dz) * inv_det;
  const double t = (k.e2x * qx + k.e2y * qy + k.e2z * qz) * inv_det;
  if (u < 0.0 || u > 1.0) return false;
  if (v < 0.0 || u + v > 1.0) return false;
  if (t < 0.0 || t > t_io) return false;
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

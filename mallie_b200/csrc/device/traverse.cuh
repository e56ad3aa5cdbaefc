// traverse.cuh -- device-side restatement of BVHAccel::Traverse and its helpers
// (bvh_accel.cc:546-844) for sm_100a.  Compiled with -fmad=false: every multiply
// and add below rounds separately, in the reference's evaluation order, so the
// hit record is bit-identical to the CPU reference.
//
// Equivalence with the reference loop (bvh_accel.cc:805-834), which pops a node,
// box-tests it against the *current* hitT, and then pushes far/near children:
//   * here both child boxes are tested when their parent is visited (they sit in
//     the parent's 128-byte PairNode).  The three slab conditions
//       (tmax > 0) && (tmin <= tmax) && (tmin <= hitT)
//     are evaluated with the same arithmetic; only the last one depends on hitT.
//   * hitT never increases, so a child that fails (tmin <= hitT) now also fails it
//     at the reference's later pop time: dropping it immediately is exact.
//   * a child that passes is either continued with immediately (near child, hitT
//     unchanged => the reference's pop-time test passes too) or pushed together
//     with its tmin; when it is popped the remaining condition (tmin <= hitT_now)
//     is re-checked -- exactly the reference's pop-time decision.
//   * near/far order is sign[axis] as in bvh_accel.cc:818-823; triangles are
//     stored in indices_ order, so "last accepted among equal t wins" is kept.
// Consequently the sequence of leaves visited, triangles tested and the number
// of box tests (1 + 2 per accepted branch) are those of the reference.
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
  bool sx, sy, sz;   // dir < 0.0 (so -0.0 counts as non-negative), bvh_accel.cc:787-790
};

__device__ __forceinline__ void ray_setup(RayD &r, double ox, double oy, double oz, double dx, double dy, double dz) {
  r.ox = ox, r.oy = oy, r.oz = oz;
  r.dx = dx, r.dy = dy, r.dz = dz;
  r.sx = dx < 0.0, r.sy = dy < 0.0, r.sz = dz < 0.0;
  r.ix = 1.0 / dx, r.iy = 1.0 / dy, r.iz = 1.0 / dz;
}

// IntersectRayAABB (bvh_accel.cc:550-593).  Ternaries, not fmin/fmax: NaNs from
// 0*inf propagate exactly as in the reference.  b = {bmin[3], bmax[3]}.
__device__ __forceinline__ bool slab_test(const double b0, const double b1, const double b2, const double b3,
                                          const double b4, const double b5, const RayD &r, const double max_t,
                                          double &tmin_out) {
  const double min_x = r.sx ? b3 : b0;
  const double min_y = r.sy ? b4 : b1;
  const double min_z = r.sz ? b5 : b2;
  const double max_x = r.sx ? b0 : b3;
  const double max_y = r.sy ? b1 : b4;
  const double max_z = r.sz ? b2 : b5;

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

// TriangleIsect (bvh_accel.cc:595-638): Moeller-Trumbore, no culling.
__device__ __forceinline__ bool tri_test(double &t_io, double &u_out, double &v_out, const double p0x,
                                         const double p0y, const double p0z, const double p1x, const double p1y,
                                         const double p1z, const double p2x, const double p2y, const double p2z,
                                         const RayD &r) {
  const double e1x = p1x - p0x, e1y = p1y - p0y, e1z = p1z - p0z;
  const double e2x = p2x - p0x, e2y = p2y - p0y, e2z = p2z - p0z;
  // p = dir x e2
  const double px = r.dy * e2z - r.dz * e2y;
  const double py = r.dz * e2x - r.dx * e2z;
  const double pz = r.dx * e2y - r.dy * e2x;
  const double det = e1x * px + e1y * py + e1z * pz;
  if (fabs(det) < MB200_TRI_EPS) return false;
  const double inv_det = 1.0 / det;
  const double sx = r.ox - p0x, sy = r.oy - p0y, sz = r.oz - p0z;
  // q = s x e1
  const double qx = sy * e1z - sz * e1y;
  const double qy = sz * e1x - sx * e1z;
  const double qz = sx * e1y - sy * e1x;
  const double u = (sx * px + sy * py + sz * pz) * inv_det;
  const double v = (qx * r.dx + qy * r.dy + qz * r.dz) * inv_det;
  const double t = (e2x * qx + e2y * qy + e2z * qz) * inv_det;
  if (u < 0.0 || u > 1.0) return false;
  if (v < 0.0 || u + v > 1.0) return false;
  if (t < 0.0 || t > t_io) return false;
  t_io = t;
  u_out = u;
  v_out = v;
  return true;
}

struct HitD {
  double t, u, v;
  uint32_t face, mat;
};

// ---- triangle record loaders ------------------------------------------------
struct TriVerts {
  double p0x, p0y, p0z, p1x, p1y, p1z, p2x, p2y, p2z;
  uint32_t face, mat;
};

template <bool F32> __device__ __forceinline__ TriVerts load_tri(const void *tris, uint32_t i);

template <> __device__ __forceinline__ TriVerts load_tri<true>(const void *tris, uint32_t i) {
  const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const TriRecordF32 *>(tris) + i);
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  TriVerts t;
  t.p0x = (double)a.x, t.p0y = (double)a.y, t.p0z = (double)a.z;
  t.face = __float_as_uint(a.w);
  t.p1x = (double)b.x, t.p1y = (double)b.y, t.p1z = (double)b.z;
  t.mat = __float_as_uint(b.w);
  t.p2x = (double)c.x, t.p2y = (double)c.y, t.p2z = (double)c.z;
  return t;
}

template <> __device__ __forceinline__ TriVerts load_tri<false>(const void *tris, uint32_t i) {
  const double2 *p = reinterpret_cast<const double2 *>(reinterpret_cast<const TriRecordF64 *>(tris) + i);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
  TriVerts t;
  t.p0x = a.x, t.p0y = a.y, t.p0z = b.x;
  t.p1x = b.y, t.p1y = c.x, t.p1z = c.y;
  t.p2x = d.x, t.p2y = d.y, t.p2z = e.x;
  const unsigned long long w = (unsigned long long)__double_as_longlong(e.y);
  t.face = (uint32_t)(w & 0xFFFFFFFFull);
  t.mat = (uint32_t)(w >> 32);
  return t;
}

// ---- traversal stack: first S entries per thread in shared memory (column
// layout: entry k of thread t at [k * blockDim.x + t] -> conflict-free 128-bit
// accesses), the rest in a per-thread local-memory array.
template <int S, int CAP> struct TravStack {
  uint4 *sm; // this thread's column base
  int stride;
  uint4 ovf[(CAP > S) ? (CAP - S) : 1];
  __device__ __forceinline__ void put(int k, double tmin, uint32_t ref, uint32_t cnt) {
    const unsigned long long tb = (unsigned long long)__double_as_longlong(tmin);
    const uint4 e = make_uint4((uint32_t)tb, (uint32_t)(tb >> 32), ref, cnt);
    if (k < S) sm[k * stride] = e;
    else ovf[k - S] = e;
  }
  __device__ __forceinline__ void get(int k, double &tmin, uint32_t &ref, uint32_t &cnt) const {
    const uint4 e = (k < S) ? sm[k * stride] : ovf[k - S];
    tmin = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
    ref = e.z;
    cnt = e.w;
  }
};

struct TravCounters {
  unsigned int nodes, tris, max_stack;
};

// Closest hit (ANYHIT = false): on entry hit.t = DBL_MAX, hit.u = hit.v = 0, face = mat = ~0
// (bvh_accel.cc:783-786); returns true iff something was hit.
// Any hit (ANYHIT = true): on entry hit.t = tmax; returns true as soon as a triangle is
// accepted with t < tmax (see kernels.cu, occlusion).
template <bool F32, int S, int CAP, bool ANYHIT, bool COUNT>
__device__ __forceinline__ bool traverse(const SceneView &sc, const RayD &r, HitD &hit, TravStack<S, CAP> &st,
                                         TravCounters &cnt) {
  if (sc.empty) return false;
  double hit_t = hit.t;
  const double tmax_any = hit.t;
  bool found = false;
  int sp = 0;

  uint32_t ref = sc.root_ref, rc = sc.root_cnt;
  {
    double tm;
    if (COUNT) cnt.nodes++;
    if (!slab_test(sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5],
                   r, hit_t, tm))
      return false;
  }

  for (;;) {
    if (rc == kBranch) {
      // ---- inner node: one 128-byte line, both children tested --------------
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
      // near = data[dirSign[axis]] (bvh_accel.cc:818-823)
      if (h0 && h1) {
        const uint32_t far_ref = sgn ? meta.x : meta.y, far_cnt = sgn ? meta.z : meta.w;
        const double far_t = sgn ? t0 : t1;
        st.put(sp++, far_t, far_ref, far_cnt);
        if (COUNT) cnt.max_stack = max(cnt.max_stack, (unsigned int)sp + 1u);
        ref = sgn ? meta.y : meta.x;
        rc = sgn ? meta.w : meta.z;
        continue;
      }
      if (h0) {
        ref = meta.x, rc = meta.z;
        continue;
      }
      if (h1) {
        ref = meta.y, rc = meta.w;
        continue;
      }
    } else {
      // ---- leaf: TestLeafNode (bvh_accel.cc:640-697) ------------------------
      if (COUNT) cnt.tris += rc;
      for (uint32_t i = 0; i < rc; i++) {
        const TriVerts tv = load_tri<F32>(sc.tris, ref + i);
        double u, v;
        if (tri_test(hit_t, u, v, tv.p0x, tv.p0y, tv.p0z, tv.p1x, tv.p1y, tv.p1z, tv.p2x, tv.p2y, tv.p2z, r)) {
          hit.t = hit_t;
          hit.u = u;
          hit.v = v;
          hit.face = tv.face;
          hit.mat = tv.mat;
          found = true;
          if (ANYHIT && hit_t < tmax_any) return true;
        }
      }
    }
    // ---- pop: the reference's pop-time (tmin <= hitT) decision ---------------
    bool got = false;
    while (sp > 0) {
      double tm;
      st.get(--sp, tm, ref, rc);
      if (tm <= hit_t) {
        got = true;
        break;
      }
    }
    if (!got) break;
  }
  return ANYHIT ? false : found;
}

} // namespace mb200

#endif

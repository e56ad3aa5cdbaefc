// kernels.cu -- sm_100a kernels of the Mallie render hot path.
//
//   K1 raygen        Camera::GenerateRay           camera.cc:222-240
//   K2 closest hit   BVHAccel::Traverse            bvh_accel.cc:773-844
//   K3 hit record    BuildIntersection             bvh_accel.cc:699-769
//   K4 occlusion     closest-hit t < tmax          (render.cc:425-426 leaves NEE empty)
//   K5 render pass   Render/PathTrace + Plane      render.cc:381-456,593-708, prim-plane.cc:8-44
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
// (no FMA contraction: see traverse.cuh).
#include "kernels.h"

#include <cstdio>

#include "traverse.cuh"

namespace mb200 {

namespace {

constexpr int kBlock = 128;   // threads per CTA
constexpr int kSmemStack = 16; // stack entries per thread kept in shared memory

__device__ __forceinline__ unsigned int lane_id() { return threadIdx.x & 31u; }

// Per-warp dynamic fetch of `per_warp` consecutive work items from a global counter
// (persistent threads: the grid is sized to the machine, not to the problem).
__device__ __forceinline__ unsigned long long warp_fetch(unsigned long long *counter, unsigned int per_warp) {
  unsigned long long base = 0;
  if (lane_id() == 0) base = atomicAdd(counter, (unsigned long long)per_warp);
  return __shfl_sync(0xFFFFFFFFu, base, 0);
}

__device__ __forceinline__ void flush_counters(const TravCounters &c, unsigned long long rays,
                                               unsigned long long *g /* [4] */) {
  unsigned long long n = c.nodes, t = c.tris, r = rays;
  unsigned int m = c.max_stack;
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_down_sync(0xFFFFFFFFu, n, o);
    t += __shfl_down_sync(0xFFFFFFFFu, t, o);
    r += __shfl_down_sync(0xFFFFFFFFu, r, o);
    m = max(m, __shfl_down_sync(0xFFFFFFFFu, m, o));
  }
  if (lane_id() == 0) {
    atomicAdd(&g[0], n);
    atomicAdd(&g[1], t);
    atomicAdd(&g[2], r);
    atomicMax(&g[3], (unsigned long long)m);
  }
}

// real3::normalize (common.h:48-57)
__device__ __forceinline__ void normalize3(double &x, double &y, double &z) {
  const double len = sqrt(x * x + y * y + z * z);
  if (fabs(len) > 1.0e-6) {
    const double inv = 1.0 / len;
    x *= inv, y *= inv, z *= inv;
  }
}

// Camera::GenerateRay (camera.cc:222-240)
__device__ __forceinline__ void generate_ray(const mb200_camera_frame &f, double u, double v, double &dx, double &dy,
                                             double &dz) {
  dx = (f.corner[0] + u * f.du[0] + v * f.dv[0]) - f.origin[0];
  dy = (f.corner[1] + u * f.du[1] + v * f.dv[1]) - f.origin[1];
  dz = (f.corner[2] + u * f.du[2] + v * f.dv[2]) - f.origin[2];
  normalize3(dx, dy, dz);
}

// ---------------------------------------------------------------------------
// K1: ray generation
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_generate_rays(const __grid_constant__ mb200_camera_frame frame,
                                                       const double *__restrict__ px, const double *__restrict__ py,
                                                       size_t n, mb200_ray *__restrict__ rays) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double dx, dy, dz;
    generate_ray(frame, px[i], py[i], dx, dy, dz);
    mb200_ray r;
    r.org[0] = frame.origin[0], r.org[1] = frame.origin[1], r.org[2] = frame.origin[2];
    r.dir[0] = dx, r.dir[1] = dy, r.dir[2] = dz;
    rays[i] = r;
  }
}

__global__ void __launch_bounds__(256) k_generate_grid(const __grid_constant__ mb200_camera_frame frame, int x0,
                                                       int y0, int w, int h, mb200_ray *__restrict__ rays) {
  const size_t n = (size_t)w * h;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = x0 + (int)(i % w), y = y0 + (int)(i / w);
    double dx, dy, dz;
    generate_ray(frame, (double)x, (double)y, dx, dy, dz);
    mb200_ray r;
    r.org[0] = frame.origin[0], r.org[1] = frame.origin[1], r.org[2] = frame.origin[2];
    r.dir[0] = dx, r.dir[1] = dy, r.dir[2] = dz;
    rays[i] = r;
  }
}

// ---------------------------------------------------------------------------
// K2 / K4: batched queries over a ray buffer (persistent warps, 32 rays per fetch)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_ray(const mb200_ray *rays, size_t i, RayD &r) {
  const double2 *p = reinterpret_cast<const double2 *>(rays + i);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  ray_setup(r, a.x, a.y, b.x, b.y, c.x, c.y);
}

template <bool F32, int CAP, bool COUNT>
__global__ void __launch_bounds__(kBlock)
    k_trace_closest(const __grid_constant__ SceneView sc, const mb200_ray *__restrict__ rays, size_t n,
                    mb200_hit *__restrict__ hits, unsigned long long *__restrict__ work,
                    unsigned long long *__restrict__ gcounters) {
  extern __shared__ uint4 smem_stack[];
  TravStack<kSmemStack, CAP> st;
  st.sm = smem_stack + threadIdx.x;
  st.stride = kBlock;
  TravCounters cnt = {0u, 0u, 0u};
  unsigned long long nrays = 0;

  for (;;) {
    const unsigned long long base = warp_fetch(work, 32u);
    if (base >= n) break;
    const size_t i = base + lane_id();
    if (i < n) {
      RayD r;
      load_ray(rays, i, r);
      HitD h;
      h.t = DBL_MAX, h.u = 0.0, h.v = 0.0, h.face = 0xFFFFFFFFu, h.mat = 0xFFFFFFFFu;
      traverse<F32, kSmemStack, CAP, false, COUNT>(sc, r, h, st, cnt);
      nrays++;
      double2 *o = reinterpret_cast<double2 *>(hits + i);
      o[0] = make_double2(h.t, h.u);
      const unsigned long long ids = ((unsigned long long)h.mat << 32) | h.face;
      o[1] = make_double2(h.v, __longlong_as_double((long long)ids));
    }
  }
  if (COUNT) flush_counters(cnt, nrays, gcounters);
}

template <bool F32, int CAP, bool COUNT>
__global__ void __launch_bounds__(kBlock)
    k_trace_occluded(const __grid_constant__ SceneView sc, const mb200_ray *__restrict__ rays,
                     const double *__restrict__ tmax, size_t n, unsigned char *__restrict__ occluded,
                     unsigned long long *__restrict__ work, unsigned long long *__restrict__ gcounters) {
  extern __shared__ uint4 smem_stack[];
  TravStack<kSmemStack, CAP> st;
  st.sm = smem_stack + threadIdx.x;
  st.stride = kBlock;
  TravCounters cnt = {0u, 0u, 0u};
  unsigned long long nrays = 0;

  for (;;) {
    const unsigned long long base = warp_fetch(work, 32u);
    if (base >= n) break;
    const size_t i = base + lane_id();
    if (i < n) {
      RayD r;
      load_ray(rays, i, r);
      HitD h;
      h.t = tmax[i], h.u = 0.0, h.v = 0.0, h.face = 0xFFFFFFFFu, h.mat = 0xFFFFFFFFu;
      const bool occ = traverse<F32, kSmemStack, CAP, true, COUNT>(sc, r, h, st, cnt);
      nrays++;
      occluded[i] = occ ? 1 : 0;
    }
  }
  if (COUNT) flush_counters(cnt, nrays, gcounters);
}

// ---------------------------------------------------------------------------
// K3: BuildIntersection (bvh_accel.cc:699-769) from a 32-byte hit record
// ---------------------------------------------------------------------------
struct IsectD {
  double px, py, pz;    // position
  double gx, gy, gz;    // geometric normal
  double nx, ny, nz;    // shading normal
  double tu, tv;        // texcoord
  uint32_t f0, f1, f2;
};

__device__ __forceinline__ void build_intersection(const SceneView &sc, const RayD &r, const HitD &h, IsectD &o) {
  const uint32_t *f = sc.faces + 3 * (size_t)h.face;
  o.f0 = __ldg(f), o.f1 = __ldg(f + 1), o.f2 = __ldg(f + 2);
  const double *v0 = sc.vertices + 3 * (size_t)o.f0;
  const double *v1 = sc.vertices + 3 * (size_t)o.f1;
  const double *v2 = sc.vertices + 3 * (size_t)o.f2;
  const double p0x = __ldg(v0), p0y = __ldg(v0 + 1), p0z = __ldg(v0 + 2);
  const double p1x = __ldg(v1), p1y = __ldg(v1 + 1), p1z = __ldg(v1 + 2);
  const double p2x = __ldg(v2), p2y = __ldg(v2 + 1), p2z = __ldg(v2 + 2);
  o.px = r.ox + h.t * r.dx;
  o.py = r.oy + h.t * r.dy;
  o.pz = r.oz + h.t * r.dz;
  const double ax = p1x - p0x, ay = p1y - p0y, az = p1z - p0z;
  const double bx = p2x - p0x, by = p2y - p0y, bz = p2z - p0z;
  double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  normalize3(nx, ny, nz);
  o.gx = nx, o.gy = ny, o.gz = nz;
  if (sc.fv_normals) {
    const double *N = sc.fv_normals + 9 * (size_t)h.face;
    const double w = 1.0 - h.u - h.v;
    o.nx = w * __ldg(N + 0) + h.u * __ldg(N + 3) + h.v * __ldg(N + 6);
    o.ny = w * __ldg(N + 1) + h.u * __ldg(N + 4) + h.v * __ldg(N + 7);
    o.nz = w * __ldg(N + 2) + h.u * __ldg(N + 5) + h.v * __ldg(N + 8);
  } else {
    o.nx = nx, o.ny = ny, o.nz = nz;
  }
  o.tu = 0.0, o.tv = 0.0;
  if (sc.fv_uvs) {
    const double *T = sc.fv_uvs + 6 * (size_t)h.face;
    const double w = 1.0 - h.u - h.v;
    o.tu = w * __ldg(T + 0) + h.u * __ldg(T + 2) + h.v * __ldg(T + 4);
    o.tv = w * __ldg(T + 1) + h.u * __ldg(T + 3) + h.v * __ldg(T + 5);
  }
}

template <bool F32, int CAP>
__global__ void __launch_bounds__(kBlock)
    k_trace_closest_full(const __grid_constant__ SceneView sc, const mb200_ray *__restrict__ rays, size_t n,
                         mb200_isect *__restrict__ isects, unsigned char *__restrict__ mask,
                         unsigned long long *__restrict__ work) {
  extern __shared__ uint4 smem_stack[];
  TravStack<kSmemStack, CAP> st;
  st.sm = smem_stack + threadIdx.x;
  st.stride = kBlock;
  TravCounters cnt = {0u, 0u, 0u};
  for (;;) {
    const unsigned long long base = warp_fetch(work, 32u);
    if (base >= n) break;
    const size_t i = base + lane_id();
    if (i < n) {
      RayD r;
      load_ray(rays, i, r);
      HitD h;
      h.t = DBL_MAX, h.u = 0.0, h.v = 0.0, h.face = 0xFFFFFFFFu, h.mat = 0xFFFFFFFFu;
      const bool hit = traverse<F32, kSmemStack, CAP, false, false>(sc, r, h, st, cnt);
      mb200_isect o;
      memset(&o, 0, sizeof(o));
      o.t = h.t, o.u = h.u, o.v = h.v, o.faceID = h.face, o.materialID = h.mat;
      if (hit) {
        IsectD d;
        build_intersection(sc, r, h, d);
        o.f0 = d.f0, o.f1 = d.f1, o.f2 = d.f2;
        o.position[0] = d.px, o.position[1] = d.py, o.position[2] = d.pz;
        o.geometricNormal[0] = d.gx, o.geometricNormal[1] = d.gy, o.geometricNormal[2] = d.gz;
        o.normal[0] = d.nx, o.normal[1] = d.ny, o.normal[2] = d.nz;
        o.texcoord[0] = d.tu, o.texcoord[1] = d.tv;
      }
      isects[i] = o;
      if (mask) mask[i] = hit ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------
// K5: one render pass (or several accumulated) -- Render + PathTrace
// ---------------------------------------------------------------------------
struct Xorshift128 { // randomreal, render.cc:137-168
  uint32_t x, y, z, w;
  __device__ __forceinline__ double next() {
    const uint32_t t = x ^ (x << 11);
    x = y, y = z, z = w;
    w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
    return w * (1.0 / 4294967296.0);
  }
};

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16, h *= 0x85ebca6bu, h ^= h >> 13, h *= 0xc2b2ae35u, h ^= h >> 16;
  return h;
}

// Replaces the per-OpenMP-thread seed table gSeed[tid] (render.cc:116-135): one
// stream per (pixel, pass), same xorshift128 generator.
__device__ __forceinline__ void rng_seed_pixel(Xorshift128 &g, uint32_t pixel, uint32_t pass) {
  const uint32_t k = mix32(pixel * 0x9e3779b9u + 0x7f4a7c15u) ^ mix32(pass * 0x85ebca6bu + 0x165667b1u);
  g.x = 123456789u ^ mix32(k + 1u);
  g.y = 362436069u ^ mix32(k + 2u);
  g.z = 521288629u ^ mix32(k + 3u);
  g.w = 88675123u ^ mix32(k + 4u);
  if ((g.x | g.y | g.z | g.w) == 0u) g.w = 88675123u;
}

// Plane::intersect (prim-plane.cc:8-44): float vn / on_d / t.
__device__ __forceinline__ bool plane_intersect(const float pl[4], const RayD &r, double &t_io, double &nx, double &ny,
                                                double &nz, uint32_t &mat) {
  double a = (double)pl[0], b = (double)pl[1], c = (double)pl[2];
  double vx = r.dx, vy = r.dy, vz = r.dz;
  normalize3(vx, vy, vz);
  const float vn = (float)(vx * a + vy * b + vz * c);
  if (fabsf(vn) > 1.1920928955078125e-7f * 1024.0f) {
    const float on_d = (float)((r.ox * a + r.oy * b + r.oz * c) + (double)pl[3]);
    const float t = -on_d / vn;
    if ((t > 0) && ((double)t < t_io)) {
      t_io = (double)t;
      normalize3(a, b, c);
      nx = a, ny = b, nz = c;
      mat = 0xFFFFFFFFu;
      return true;
    }
  }
  return false;
}

// GenerateBasis + SampleDiffuseIS (render.cc:271-339)
__device__ __forceinline__ void sample_diffuse(Xorshift128 &rng, double nx, double ny, double nz, double &ox,
                                               double &oy, double &oz) {
  int index = -1;
  double minval = 1.0e+6;
  {
    double val = (double)fabsf((float)nx);
    if (val < minval) minval = val, index = 0;
    val = (double)fabsf((float)ny);
    if (val < minval) minval = val, index = 1;
    val = (double)fabsf((float)nz);
    if (val < minval) minval = val, index = 2;
  }
  double tx, ty, tz;
  if (index == 0) tx = 0.0, ty = -nz, tz = ny;
  else if (index == 1) tx = -nz, ty = 0.0, tz = nx;
  else tx = -ny, ty = nx, tz = 0.0;
  normalize3(tx, ty, tz);
  double bx = ty * nz - tz * ny, by = tz * nx - tx * nz, bz = tx * ny - ty * nx;
  normalize3(bx, by, bz);
  const double theta = acos(sqrt(1.0 - rng.next()));
  const double phi = 2.0 * 3.14159265358979323846 * rng.next();
  const double ct = cos(theta), st = sin(theta), cp = cos(phi), sp = sin(phi);
  ox = ((tx * cp) * st + (bx * sp) * st) + nx * ct;
  oy = ((ty * cp) * st + (by * sp) * st) + ny * ct;
  oz = ((tz * cp) * st + (bz * sp) * st) + nz * ct;
}

struct RenderCounters {
  unsigned int primary, bounce, shadow, zombie;
};

constexpr double kRenderEPS = 1.0e-3; // render.cc:51
constexpr double kFar = 1.0e+30;      // render.cc:50

// One sample of one pixel.  PathTrace (render.cc:381-456) with the post-escape
// "zombie" segments resolved in closed form (they cannot hit),
// or the primary+shadow shader.
template <bool F32, int CAP>
__device__ __forceinline__ void shade_sample(const SceneView &sc, const mb200_render_params &p, int px, int py,
                                             uint32_t pass, TravStack<kSmemStack, CAP> &st, RenderCounters &rc,
                                             double &out_r, double &out_g, double &out_b) {
  TravCounters tc = {0u, 0u, 0u};
  Xorshift128 rng;
  rng_seed_pixel(rng, (uint32_t)((size_t)py * p.width + px), pass);
  double fu = (double)px, fv = (double)py;
  if (p.jitter) {
    const float ju = (float)(rng.next() - 0.5);
    const float jv = (float)(rng.next() - 0.5);
    fu = (double)((float)px + ju); // int + float is a float add (render.cc:391)
    fv = (double)((float)py + jv);
  }
  RayD r;
  {
    double dx, dy, dz;
    generate_ray(p.frame, fu, fv, dx, dy, dz);
    ray_setup(r, p.frame.origin[0], p.frame.origin[1], p.frame.origin[2], dx, dy, dz);
  }
  out_r = out_g = out_b = 0.0;

  double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;
  uint32_t cur_mat = 0; // Intersection::materialID is zero-initialised in the oracle harness
  bool escaped = false;
  const unsigned int max_len = (unsigned int)p.max_path_length;

  for (unsigned int len = 1;; ++len) {
    bool hit = false;
    HitD h;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    if (!escaped) {
      h.t = DBL_MAX, h.u = 0.0, h.v = 0.0, h.face = 0xFFFFFFFFu, h.mat = cur_mat;
      if (len == 1) rc.primary++;
      else rc.bounce++;
      hit = traverse<F32, kSmemStack, CAP, false, false>(sc, r, h, st, tc);
      if (hit) {
        IsectD d;
        build_intersection(sc, r, h, d);
        nx = d.nx, ny = d.ny, nz = d.nz;
      }
      if (p.use_plane) hit |= plane_intersect(p.plane, r, h.t, nx, ny, nz, h.mat);
      cur_mat = h.mat;
    } else {
      rc.zombie++;
    }

    if (p.shader != MB200_SHADER_PATHTRACE) {
      if (!hit) return;
      if (p.shader == MB200_SHADER_PRIMARY_ONLY) {
        out_r = out_g = out_b = 1.0;
        return;
      }
      // primary + shadow: the NEE block the reference leaves empty.
      const double hx = r.ox + h.t * r.dx, hy = r.oy + h.t * r.dy, hz = r.oz + h.t * r.dz;
      if ((nx * (-r.dx) + ny * (-r.dy) + nz * (-r.dz)) < 0.0) nx = -nx, ny = -ny, nz = -nz;
      double lx = p.light[0] - hx, ly = p.light[1] - hy, lz = p.light[2] - hz;
      const double dist = sqrt(lx * lx + ly * ly + lz * lz);
      normalize3(lx, ly, lz);
      RayD sr;
      ray_setup(sr, hx + lx * kRenderEPS, hy + ly * kRenderEPS, hz + lz * kRenderEPS, lx, ly, lz);
      HitD sh;
      sh.t = dist - kRenderEPS, sh.u = 0.0, sh.v = 0.0, sh.face = 0xFFFFFFFFu, sh.mat = 0xFFFFFFFFu;
      rc.shadow++;
      const bool occ = traverse<F32, kSmemStack, CAP, true, false>(sc, sr, sh, st, tc);
      const double ndotl = nx * lx + ny * ly + nz * lz;
      if (occ || !(ndotl > 0.0)) return;
      const double kd = (h.mat != 0xFFFFFFFFu) ? 0.5 : 1.0;
      out_r = out_g = out_b = kd * ndotl;
      return;
    }

    if (!hit) {
      if (len < 2) return; // kMinPathLength: eye ray escaped
      const double l = (double)len;
      out_r += thr_r * 0.5 / l, out_g += thr_g * 0.5 / l, out_b += thr_b * 0.5 / l;
      escaped = true;
    }
    if (len >= max_len) return;

    if (escaped) {
      if (cur_mat != 0xFFFFFFFFu) thr_r *= 0.5, thr_g *= 0.5, thr_b *= 0.5;
      continue;
    }
    const double hx = r.ox + h.t * r.dx, hy = r.oy + h.t * r.dy, hz = r.oz + h.t * r.dz;
    (void)rng.next(); // drawn, unused (render.cc:430)
    if ((nx * (-r.dx) + ny * (-r.dy) + nz * (-r.dz)) < 0.0) nx = -nx, ny = -ny, nz = -nz;
    double sx, sy, sz;
    sample_diffuse(rng, nx, ny, nz, sx, sy, sz);
    if (cur_mat != 0xFFFFFFFFu) thr_r *= 0.5, thr_g *= 0.5, thr_b *= 0.5; // default Material::diffuse (scene.h:58-65)
    ray_setup(r, hx + sx * kRenderEPS, hy + sy * kRenderEPS, hz + sz * kRenderEPS, sx, sy, sz);
  }
}

// Rows owned by band `index` of `count` when `rows` scanlines are cut into bands of `band_rows`.
__host__ __device__ inline int band_local_rows(int rows, int band_rows, int count, int index) {
  const int nbands = (rows + band_rows - 1) / band_rows;
  int local = 0;
  for (int b = index; b < nbands; b += count) {
    const int lo = b * band_rows, hi = lo + band_rows < rows ? lo + band_rows : rows;
    local += hi - lo;
  }
  return local;
}

// Persistent warps; each fetch is one 8x4 pixel tile of the render rectangle (coherent
// primary rays per warp).  num_passes samples per pixel are taken back to back.
// mode 0: image = last pass, count += passes (one pass: render.cc:673-679)
// mode 1: image += passes, count += passes   (AccumImage, main_sdl.cc:138-143)
// mode 2: image = sum of passes, count = passes (fresh frame; nothing read)
template <bool F32, int CAP>
__global__ void __launch_bounds__(kBlock)
    k_render(const __grid_constant__ SceneView sc, const __grid_constant__ mb200_render_params p, int num_passes,
             int mode, float *__restrict__ image, int *__restrict__ count,
             unsigned long long *__restrict__ work, unsigned long long *__restrict__ gstats) {
  extern __shared__ uint4 smem_stack[];
  TravStack<kSmemStack, CAP> st;
  st.sm = smem_stack + threadIdx.x;
  st.stride = kBlock;
  RenderCounters rc = {0u, 0u, 0u, 0u};

  // rows this call owns: all of [y0,y1), or every band_count-th band of band_rows scanlines
  const int rows_total = p.y1 - p.y0;
  int rows_local = rows_total;
  if (p.band_rows > 0) rows_local = band_local_rows(rows_total, p.band_rows, p.band_count, p.band_index);
  const int tw = (p.x1 - p.x0 + 7) >> 3, th = (rows_local + 3) >> 2;
  const unsigned long long ntiles = (unsigned long long)tw * th;
  for (;;) {
    const unsigned long long tile = warp_fetch(work, 1u);
    if (tile >= ntiles) break;
    const int tx = (int)(tile % tw), ty = (int)(tile / tw);
    const int x = p.x0 + tx * 8 + (int)(lane_id() & 7u);
    const int rl = ty * 4 + (int)(lane_id() >> 3); // row among the rows this call owns
    int y = p.y0 + rl;
    if (p.band_rows > 0) y = p.y0 + ((rl / p.band_rows) * p.band_count + p.band_index) * p.band_rows + rl % p.band_rows;
    if (x < p.x1 && rl < rows_local) {
      const size_t pix = (p.band_rows > 0 && p.band_compact) ? ((size_t)rl * p.width + x) : ((size_t)y * p.width + x);
      float ar = 0.f, ag = 0.f, ab = 0.f;
      if (mode == 1) ar = image[3 * pix + 0], ag = image[3 * pix + 1], ab = image[3 * pix + 2];
      for (int s = 0; s < num_passes; s++) {
        double r, g, b;
        shade_sample<F32, CAP>(sc, p, x, y, p.pass + (uint32_t)s, st, rc, r, g, b);
        if (mode != 0) ar += (float)r, ag += (float)g, ab += (float)b; // AccumImage: float += float
        else ar = (float)r, ag = (float)g, ab = (float)b;
      }
      image[3 * pix + 0] = ar, image[3 * pix + 1] = ag, image[3 * pix + 2] = ab;
      if (mode == 2) count[pix] = num_passes;
      else count[pix] += num_passes;
    }
  }
  unsigned long long a = rc.primary, b = rc.bounce, c = rc.shadow, d = rc.zombie;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xFFFFFFFFu, a, o);
    b += __shfl_down_sync(0xFFFFFFFFu, b, o);
    c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    d += __shfl_down_sync(0xFFFFFFFFu, d, o);
  }
  if (lane_id() == 0) {
    atomicAdd(&gstats[0], a);
    atomicAdd(&gstats[1], b);
    atomicAdd(&gstats[2], c);
    atomicAdd(&gstats[3], d);
  }
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
int g_num_sms = 0;
int g_launches = 0;

int persistent_grid(const void *kernel, size_t smem) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem);
  if (per_sm < 1) per_sm = 1;
  return g_num_sms * per_sm; // a whole number of CTAs per SM: one persistent wave
}

constexpr size_t kStackSmem = (size_t)kSmemStack * kBlock * sizeof(uint4);

template <typename K> cudaError_t prepare(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStackSmem);
}

} // namespace

int launches_issued() { return g_launches; }

int band_rows_owned(int rows, int band_rows, int count, int index) {
  return band_local_rows(rows, band_rows, count, index);
}

template <bool F32, int CAP>
static cudaError_t do_trace_closest(const SceneView &sc, const mb200_ray *rays, size_t n, mb200_hit *hits,
                                    unsigned long long *work, unsigned long long *counters, cudaStream_t s) {
  cudaError_t e;
  if (counters) {
    auto k = k_trace_closest<F32, CAP, true>;
    if ((e = prepare(k)) != cudaSuccess) return e;
    k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, rays, n, hits, work, counters);
  } else {
    auto k = k_trace_closest<F32, CAP, false>;
    if ((e = prepare(k)) != cudaSuccess) return e;
    k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, rays, n, hits, work, counters);
  }
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_trace_closest(const SceneView &sc, int stack_cap, const mb200_ray *rays, size_t n,
                                 mb200_hit *hits, unsigned long long *work, unsigned long long *counters,
                                 cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (sc.tri_f32) {
    return stack_cap <= 64 ? do_trace_closest<true, 64>(sc, rays, n, hits, work, counters, s)
                           : do_trace_closest<true, 512>(sc, rays, n, hits, work, counters, s);
  }
  return stack_cap <= 64 ? do_trace_closest<false, 64>(sc, rays, n, hits, work, counters, s)
                         : do_trace_closest<false, 512>(sc, rays, n, hits, work, counters, s);
}

template <bool F32, int CAP>
static cudaError_t do_trace_occluded(const SceneView &sc, const mb200_ray *rays, const double *tmax, size_t n,
                                     unsigned char *occ, unsigned long long *work, unsigned long long *counters,
                                     cudaStream_t s) {
  cudaError_t e;
  if (counters) {
    auto k = k_trace_occluded<F32, CAP, true>;
    if ((e = prepare(k)) != cudaSuccess) return e;
    k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, rays, tmax, n, occ, work,
                                                                               counters);
  } else {
    auto k = k_trace_occluded<F32, CAP, false>;
    if ((e = prepare(k)) != cudaSuccess) return e;
    k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, rays, tmax, n, occ, work,
                                                                               counters);
  }
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_trace_occluded(const SceneView &sc, int stack_cap, const mb200_ray *rays, const double *tmax,
                                  size_t n, unsigned char *occ, unsigned long long *work,
                                  unsigned long long *counters, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (sc.tri_f32) {
    return stack_cap <= 64 ? do_trace_occluded<true, 64>(sc, rays, tmax, n, occ, work, counters, s)
                           : do_trace_occluded<true, 512>(sc, rays, tmax, n, occ, work, counters, s);
  }
  return stack_cap <= 64 ? do_trace_occluded<false, 64>(sc, rays, tmax, n, occ, work, counters, s)
                         : do_trace_occluded<false, 512>(sc, rays, tmax, n, occ, work, counters, s);
}

template <bool F32, int CAP>
static cudaError_t do_trace_full(const SceneView &sc, const mb200_ray *rays, size_t n, mb200_isect *isects,
                                 unsigned char *mask, unsigned long long *work, cudaStream_t s) {
  auto k = k_trace_closest_full<F32, CAP>;
  cudaError_t e = prepare(k);
  if (e != cudaSuccess) return e;
  k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, rays, n, isects, mask, work);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_trace_closest_full(const SceneView &sc, int stack_cap, const mb200_ray *rays, size_t n,
                                      mb200_isect *isects, unsigned char *mask, unsigned long long *work,
                                      cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (sc.tri_f32) {
    return stack_cap <= 64 ? do_trace_full<true, 64>(sc, rays, n, isects, mask, work, s)
                           : do_trace_full<true, 512>(sc, rays, n, isects, mask, work, s);
  }
  return stack_cap <= 64 ? do_trace_full<false, 64>(sc, rays, n, isects, mask, work, s)
                         : do_trace_full<false, 512>(sc, rays, n, isects, mask, work, s);
}

template <bool F32, int CAP>
static cudaError_t do_render(const SceneView &sc, const mb200_render_params &p, int num_passes, int accumulate,
                             float *image, int *count, unsigned long long *work, unsigned long long *stats,
                             cudaStream_t s) {
  auto k = k_render<F32, CAP>;
  cudaError_t e = prepare(k);
  if (e != cudaSuccess) return e;
  k<<<persistent_grid((const void *)k, kStackSmem), kBlock, kStackSmem, s>>>(sc, p, num_passes, accumulate, image,
                                                                             count, work, stats);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_render(const SceneView &sc, int stack_cap, const mb200_render_params &p, int num_passes,
                          int accumulate, float *image, int *count, unsigned long long *work,
                          unsigned long long *stats, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (sc.tri_f32) {
    return stack_cap <= 64 ? do_render<true, 64>(sc, p, num_passes, accumulate, image, count, work, stats, s)
                           : do_render<true, 512>(sc, p, num_passes, accumulate, image, count, work, stats, s);
  }
  return stack_cap <= 64 ? do_render<false, 64>(sc, p, num_passes, accumulate, image, count, work, stats, s)
                         : do_render<false, 512>(sc, p, num_passes, accumulate, image, count, work, stats, s);
}

cudaError_t launch_generate_rays(const mb200_camera_frame &f, const double *px, const double *py, size_t n,
                                 mb200_ray *rays, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_generate_rays<<<grid, 256, 0, s>>>(f, px, py, n, rays);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_generate_grid(const mb200_camera_frame &f, int x0, int y0, int w, int h, mb200_ray *rays,
                                 cudaStream_t s) {
  const size_t n = (size_t)w * h;
  if (n == 0) return cudaSuccess;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_generate_grid<<<grid, 256, 0, s>>>(f, x0, y0, w, h, rays);
  g_launches++;
  return cudaGetLastError();
}

} // namespace mb200

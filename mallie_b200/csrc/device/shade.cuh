// shade.cuh -- device-side restatement of the per-sample arithmetic of mallie::Render / PathTrace that is
// not traversal: camera rays, the xorshift128 RNG, the debug plane, BuildIntersection, the cosine sampler,
// and the pixel <-> work-item mapping of the frame kernels.  -fmad=false; operation order as the reference.
//
//   Camera::GenerateRay            camera.cc:222-240
//   randomreal                     render.cc:137-168
//   Plane::intersect               prim-plane.cc:8-44
//   BuildIntersection              bvh_accel.cc:699-769
//   GenerateBasis/SampleDiffuseIS  render.cc:271-339
#ifndef MALLIE_B200_SHADE_CUH_
#define MALLIE_B200_SHADE_CUH_

#include "mathd.cuh"
#include "layout.h"
#include "mallie_b200.h"
#include "traverse.cuh"

namespace mb200 {

constexpr double kRenderEPS = 1.0e-3; // render.cc:51

// real3::normalize (common.h:48-57): only vectors longer than 1e-6 are rescaled
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

// Camera::GenerateEnvRay (camera.cc:242-257) / GenerateStereoEnvRay (camera.cc:259-329)
__device__ __forceinline__ void generate_env_ray(const double origin[3], int width, int height, double u, double v,
                                                 bool stereo, double &ox, double &oy, double &oz, double &dx,
                                                 double &dy, double &dz) {
  const double kPi = 3.14159265358979323846;
  ox = origin[0], oy = origin[1], oz = origin[2];
  const double phi = 2.0 * kPi * (u / (double)width);
  // sin / cos rounded once from double-double values (mathd.cuh): what the host's libm returns in ~99.9 % of its results
  double st, ct, sp, cp;
  mathd::sincos_rn(phi, sp, cp);
  if (!stereo) {
    const double theta = kPi * (v / (double)height);
    mathd::sincos_rn(theta, st, ct);
    dx = st * cp;
    dy = ct;
    dz = st * sp;
    return;
  }
  const bool left = v < (double)(height >> 1);
  const double r = 0.5; // focal_length = 4.0
  const double theta = kPi * fmod(2.0 * v / (double)height, 1.0);
  mathd::sincos_rn(theta, st, ct);
  const double ex = st * cp, ey = ct, ez = st * sp;
  double px = left ? -ez : ez, py = 0.0, pz = left ? ex : -ex;
  normalize3(px, py, pz);
  ox += px * r, oy += py * r, oz += pz * r;
  // psi = atan2(r, focal_length) = atan2(0.5, 4.0) is a constant of the reference (camera.cc:308): psi, cos(psi) and
  // sin(psi) below are the correctly rounded values (= glibc's); the left eye uses -psi, cos even, sin odd
  const double cpsi = 0.9922778767136676, spsi_abs = 0.12403473458920845;
  const double spsi = left ? -spsi_abs : spsi_abs;
  dx = ex * cpsi - ez * spsi;
  dy = ey;
  dz = ex * spsi + ez * cpsi;
  normalize3(dx, dy, dz);
}

// The camera ray of sample `pass` of pixel (px, py): PathTrace's / PathTraceEnv's prologue
// (render.cc:386-393, 524-533).  Leaves rng positioned after the two jitter draws.
// PINHOLE_ONLY: the caller guarantees p.camera_mode == MB200_CAMERA_PINHOLE (the panorama code is not compiled in).
template <bool PINHOLE_ONLY = false>
__device__ __forceinline__ void camera_sample(const mb200_render_params &p, int px, int py, uint32_t pass,
                                              Xorshift128 &rng, double &ox, double &oy, double &oz, double &dx,
                                              double &dy, double &dz) {
  rng_seed_pixel(rng, (uint32_t)((size_t)py * p.width + px), pass);
  double fu = (double)px, fv = (double)py;
  if (p.jitter) {
    const float ju = (float)(rng.next() - 0.5);
    const float jv = (float)(rng.next() - 0.5);
    fu = (double)((float)px + ju); // int + float is a float add (render.cc:391)
    fv = (double)((float)py + jv);
  }
  if (PINHOLE_ONLY || p.camera_mode == MB200_CAMERA_PINHOLE) {
    ox = p.frame.origin[0], oy = p.frame.origin[1], oz = p.frame.origin[2];
    generate_ray(p.frame, fu, fv, dx, dy, dz);
  } else {
    generate_env_ray(p.frame.origin, p.width, p.height, fu, fv, p.camera_mode == MB200_CAMERA_ENV_STEREO, ox, oy, oz,
                     dx, dy, dz);
  }
}

// Plane::intersect (prim-plane.cc:8-44): float vn / on_d / t.
__device__ __forceinline__ bool plane_intersect(const float pl[4], double ox, double oy, double oz, double dx,
                                                double dy, double dz, double &t_io, double &nx, double &ny,
                                                double &nz, uint32_t &mat) {
  double a = (double)pl[0], b = (double)pl[1], c = (double)pl[2];
  double vx = dx, vy = dy, vz = dz;
  normalize3(vx, vy, vz);
  const float vn = (float)(vx * a + vy * b + vz * c);
  if (fabsf(vn) > 1.1920928955078125e-7f * 1024.0f) {
    const float on_d = (float)((ox * a + oy * b + oz * c) + (double)pl[3]);
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

// BuildIntersection (bvh_accel.cc:699-769) from a hit record and the ray it belongs to.
struct IsectD {
  double px, py, pz; // position
  double gx, gy, gz; // geometric normal
  double nx, ny, nz; // shading normal
  double tu, tv;     // texcoord
  uint32_t f0, f1, f2;
};

__device__ __forceinline__ void build_intersection(const SceneView &sc, double ox, double oy, double oz, double dx,
                                                   double dy, double dz, double t, double u, double v, uint32_t face,
                                                   IsectD &o) {
  const uint32_t *f = sc.faces + 3 * (size_t)face;
  o.f0 = __ldg(f), o.f1 = __ldg(f + 1), o.f2 = __ldg(f + 2);
  const double *v0 = sc.vertices + 3 * (size_t)o.f0;
  const double *v1 = sc.vertices + 3 * (size_t)o.f1;
  const double *v2 = sc.vertices + 3 * (size_t)o.f2;
  const double p0x = __ldg(v0), p0y = __ldg(v0 + 1), p0z = __ldg(v0 + 2);
  const double p1x = __ldg(v1), p1y = __ldg(v1 + 1), p1z = __ldg(v1 + 2);
  const double p2x = __ldg(v2), p2y = __ldg(v2 + 1), p2z = __ldg(v2 + 2);
  o.px = ox + t * dx;
  o.py = oy + t * dy;
  o.pz = oz + t * dz;
  const double ax = p1x - p0x, ay = p1y - p0y, az = p1z - p0z;
  const double bx = p2x - p0x, by = p2y - p0y, bz = p2z - p0z;
  double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  normalize3(nx, ny, nz);
  o.gx = nx, o.gy = ny, o.gz = nz;
  if (sc.fv_normals) {
    const double *N = sc.fv_normals + 9 * (size_t)face;
    const double w = 1.0 - u - v;
    o.nx = w * __ldg(N + 0) + u * __ldg(N + 3) + v * __ldg(N + 6);
    o.ny = w * __ldg(N + 1) + u * __ldg(N + 4) + v * __ldg(N + 7);
    o.nz = w * __ldg(N + 2) + u * __ldg(N + 5) + v * __ldg(N + 8);
  } else {
    o.nx = nx, o.ny = ny, o.nz = nz;
  }
  o.tu = 0.0, o.tv = 0.0;
  if (sc.fv_uvs) {
    const double *T = sc.fv_uvs + 6 * (size_t)face;
    const double w = 1.0 - u - v;
    o.tu = w * __ldg(T + 0) + u * __ldg(T + 2) + v * __ldg(T + 4);
    o.tv = w * __ldg(T + 1) + u * __ldg(T + 3) + v * __ldg(T + 5);
  }
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

// ---- work items of a frame ---------------------------------------------------------------------------
// The pixels a render call samples (its rectangle, or its row bands of it; every step-th pixel of every
// step-th row when Render()'s step > 1) are cut into 8x4 tiles; one work
// item is one sample of one pixel, numbered   item = (tile * passes + pass) * 32 + lane   so that 32
// consecutive items are one tile (coherent rays for a warp) and the samples of a tile are adjacent (the
// same BVH region stays in L1/L2).  Items of lanes that fall outside the rectangle are "invalid".
struct FrameMap {
  int x0, x1, y0;       // render rectangle (columns, first row)
  int rows_local;       // rows this call owns
  int tiles_x;          // tiles per tile-row
  int width;            // image width (pixel index = y * width + x)
  int band_rows, band_count, band_index, compact;
  uint32_t passes;      // samples per pixel in this batch
  uint32_t pass0;       // first pass index of the batch
  int step;             // Render()'s step (>= 1): sampled pixels are step apart in x and y
  int y1;               // end row of the rectangle (block fill clipping)
  // exact division by the launch constants passes / tiles_x as one 64-bit high multiply (0 means d == 1)
  unsigned long long magic_passes, magic_tiles_x;
  // Longest-rays-first scheduling (kernels.h: FramePipe).  order (nullable = identity): the tile that work-item
  // slot k stands for, tiles that held long rays in the previous frame first; hot (nullable): per tile, set
  // by the traversal kernels when one of the tile's rays needed more than kHotSteps steps.
  const uint32_t *order;
  unsigned char *hot;
  uint32_t hot_steps;
  // First tile of the batch: a frame is cut into batches by tile rows (each batch holds ALL passes of its rows, so its
  // pixels are final when it is resolved), or -- when one tile row times the passes would not fit a batch -- by passes.
  uint32_t tile0;
};

constexpr uint32_t kHotSteps = 160; // mean ray ~35 steps, p99 ~130, silhouette rays up to ~430 (1 M-triangle sphere)

// M = ceil(2^64 / d) = (2^64 + e) / d with 0 <= e < d.  For n < 2^32: n * M / 2^64 = n / d + n * e / (d * 2^64),
// and n * e < 2^64, so the extra term is below 1 / d and floor(n * M / 2^64) == floor(n / d): exact.
__host__ __device__ inline unsigned long long div_magic(uint32_t d) {
  return d <= 1u ? 0ull : (0xFFFFFFFFFFFFFFFFull / d) + 1ull;
}

__host__ __device__ inline int band_local_rows(int rows, int band_rows, int count, int index) {
  const int nbands = (rows + band_rows - 1) / band_rows;
  int local = 0;
  for (int b = index; b < nbands; b += count) {
    const int lo = b * band_rows, hi = lo + band_rows < rows ? lo + band_rows : rows;
    local += hi - lo;
  }
  return local;
}

__host__ inline FrameMap make_frame_map(const mb200_render_params &p, uint32_t pass0, uint32_t passes) {
  FrameMap m;
  m.x0 = p.x0, m.x1 = p.x1, m.y0 = p.y0;
  m.step = p.pixel_step > 1 ? p.pixel_step : 1;
  m.y1 = p.y1;
  m.rows_local = (p.y1 - p.y0 + m.step - 1) / m.step; // sampled rows
  if (p.band_rows > 0) m.rows_local = band_local_rows(p.y1 - p.y0, p.band_rows, p.band_count, p.band_index);
  m.tiles_x = ((p.x1 - p.x0 + m.step - 1) / m.step + 7) >> 3;
  m.width = p.width;
  m.band_rows = p.band_rows, m.band_count = p.band_count, m.band_index = p.band_index, m.compact = p.band_compact;
  m.passes = passes, m.pass0 = pass0;
  m.magic_passes = div_magic(passes), m.magic_tiles_x = div_magic((uint32_t)m.tiles_x);
  m.order = nullptr, m.hot = nullptr, m.hot_steps = kHotSteps;
  m.tile0 = 0u;
  return m;
}

__host__ inline size_t frame_map_tiles(const FrameMap &m) {
  return (size_t)m.tiles_x * (size_t)((m.rows_local + 3) >> 2);
}

// item -> pixel; returns false for padding lanes.  rl = row among the rows this call owns.
__device__ __forceinline__ bool item_pixel(const FrameMap &m, uint32_t item, int &x, int &y, int &rl, uint32_t &pass) {
  const uint32_t lane = item & 31u, g = item >> 5;
  const uint32_t slot = m.magic_passes ? (uint32_t)__umul64hi((unsigned long long)g, m.magic_passes) : g;
  pass = m.pass0 + (g - slot * m.passes);
  const uint32_t tile = m.order ? __ldg(m.order + m.tile0 + slot) : m.tile0 + slot;
  const uint32_t tyu = m.magic_tiles_x ? (uint32_t)__umul64hi((unsigned long long)tile, m.magic_tiles_x) : tile;
  const int ty = (int)tyu, tx = (int)(tile - tyu * (uint32_t)m.tiles_x);
  x = m.x0 + (tx * 8 + (int)(lane & 7u)) * m.step;
  rl = ty * 4 + (int)(lane >> 3);
  y = m.y0 + rl * m.step;
  if (m.band_rows > 0) y = m.y0 + ((rl / m.band_rows) * m.band_count + m.band_index) * m.band_rows + rl % m.band_rows;
  return x < m.x1 && rl < m.rows_local;
}

// a ray of work item `item` turned out to be long: remember its tile for the next frame's schedule
__device__ __forceinline__ void mark_hot_tile(const FrameMap &m, uint32_t item) {
  if (!m.hot) return;
  const uint32_t g = item >> 5;
  const uint32_t slot = m.magic_passes ? (uint32_t)__umul64hi((unsigned long long)g, m.magic_passes) : g;
  m.hot[m.order ? __ldg(m.order + m.tile0 + slot) : m.tile0 + slot] = 1;
}

// where a pixel of this call lives in the caller's image / count buffers
__device__ __forceinline__ size_t pixel_slot(const FrameMap &m, int x, int y, int rl) {
  return (m.band_rows > 0 && m.compact) ? ((size_t)rl * m.width + x) : ((size_t)y * m.width + x);
}

} // namespace mb200

#endif

// camera.cc -- host side of the camera: Camera::BuildCameraFrame (camera.cc:40-220)
// with its helpers Matrix::LookAt / Inverse / Mult / MultV (matrix.cc:42-216) and
// build_rotmatrix (trackball.cc:272-292).  Runs once per pass on the host; the
// per-pixel Camera::GenerateRay (camera.cc:222-240) runs in the raygen kernel.
//
// The frame must be bit-identical to the reference's because every primary ray is
// derived from it, so the floating-point evaluation order is kept (including the
// reference's quirks: vector lengths truncated to float in camera.cc's normalize,
// focal length through single-precision tanf, `up` reset to the caller's vector).
#include <cmath>
#include <cstring>

#include "mallie_api.h"

namespace {

typedef double Mat4[4][4];

inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

inline void cross3(double c[3], const double a[3], const double b[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// vlength (camera.cc:22-28, matrix.cc:18-24)
inline double length3(const double v[3]) {
  const double l2 = dot3(v, v);
  return (std::fabs(l2) > 1.0e-30) ? std::sqrt(l2) : 0.0;
}

// matrix.cc:26-34 keeps the length in double ...
inline void normalize_d(double v[3]) {
  const double len = length3(v);
  if (std::fabs(len) > 1.0e-30) {
    const double inv = 1.0 / len;
    v[0] *= inv, v[1] *= inv, v[2] *= inv;
  }
}

// ... camera.cc:30-38 narrows it to float first.
inline void normalize_f(double v[3]) {
  const float len = (float)length3(v);
  if (std::fabs(len) > 1.0e-30) {
    const double inv = 1.0 / len;
    v[0] *= inv, v[1] *= inv, v[2] *= inv;
  }
}

// Matrix::LookAt (matrix.cc:42-99): rows u, v, -look, eye.
void look_at(Mat4 m, const double eye[3], const double target[3], const double up[3]) {
  double look[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
  double u[3], v[3];
  normalize_d(look);
  cross3(u, look, up);
  normalize_d(u);
  cross3(v, u, look);
  normalize_d(v);
  for (int c = 0; c < 3; c++) m[0][c] = u[c], m[1][c] = v[c], m[2][c] = -look[c], m[3][c] = eye[c];
  m[0][3] = m[1][3] = m[2][3] = 0.0;
  m[3][3] = 1.0;
}

// Matrix::Inverse (matrix.cc:101-193) is Cramer's rule on the transposed matrix with
// twelve 2x2 "pair" products reused across cofactors.  Written here as tables so the
// order of every multiply/add is explicit: each output is
//   (p[a0]*s[b0] + p[a1]*s[b1] + p[a2]*s[b2]) - (p[a3]*s[b3] + p[a4]*s[b4] + p[a5]*s[b5]).
struct Cof {
  unsigned char a[6], b[6];
};
const unsigned char kPairsHi[12][2] = {{10, 15}, {11, 14}, {9, 15}, {11, 13}, {9, 14}, {10, 13},
                                       {8, 15},  {11, 12}, {8, 14}, {10, 12}, {8, 13}, {9, 12}};
const unsigned char kPairsLo[12][2] = {{2, 7}, {3, 6}, {1, 7}, {3, 5}, {1, 6}, {2, 5},
                                       {0, 7}, {3, 4}, {0, 6}, {2, 4}, {0, 5}, {1, 4}};
const Cof kCofHi[8] = {
    {{0, 3, 4, 1, 2, 5}, {5, 6, 7, 5, 6, 7}},   {{1, 6, 9, 0, 7, 8}, {4, 6, 7, 4, 6, 7}},
    {{2, 7, 10, 3, 6, 11}, {4, 5, 7, 4, 5, 7}}, {{5, 8, 11, 4, 9, 10}, {4, 5, 6, 4, 5, 6}},
    {{1, 2, 5, 0, 3, 4}, {1, 2, 3, 1, 2, 3}},   {{0, 7, 8, 1, 6, 9}, {0, 2, 3, 0, 2, 3}},
    {{3, 6, 11, 2, 7, 10}, {0, 1, 3, 0, 1, 3}}, {{4, 9, 10, 5, 8, 11}, {0, 1, 2, 0, 1, 2}}};
const Cof kCofLo[8] = {
    {{0, 3, 4, 1, 2, 5}, {13, 14, 15, 13, 14, 15}},   {{1, 6, 9, 0, 7, 8}, {12, 14, 15, 12, 14, 15}},
    {{2, 7, 10, 3, 6, 11}, {12, 13, 15, 12, 13, 15}}, {{5, 8, 11, 4, 9, 10}, {12, 13, 14, 12, 13, 14}},
    {{2, 5, 1, 4, 0, 3}, {10, 11, 9, 11, 9, 10}},     {{8, 0, 7, 6, 9, 1}, {11, 8, 10, 10, 11, 8}},
    {{6, 11, 3, 10, 2, 7}, {9, 11, 8, 11, 8, 9}},     {{10, 4, 9, 8, 11, 5}, {10, 8, 9, 9, 0, 8}}};
//                                                      note the s[0] in the last row: matrix.cc:179

void invert(Mat4 m) {
  double s[16], p[12];
  for (int i = 0; i < 4; i++) s[i] = m[i][0], s[i + 4] = m[i][1], s[i + 8] = m[i][2], s[i + 12] = m[i][3];
  double *out = &m[0][0];
  for (int half = 0; half < 2; half++) {
    const unsigned char(*pairs)[2] = half ? kPairsLo : kPairsHi;
    const Cof *cof = half ? kCofLo : kCofHi;
    for (int k = 0; k < 12; k++) p[k] = s[pairs[k][0]] * s[pairs[k][1]];
    for (int k = 0; k < 8; k++) {
      const Cof &c = cof[k];
      double pos = p[c.a[0]] * s[c.b[0]] + p[c.a[1]] * s[c.b[1]] + p[c.a[2]] * s[c.b[2]];
      pos -= p[c.a[3]] * s[c.b[3]] + p[c.a[4]] * s[c.b[4]] + p[c.a[5]] * s[c.b[5]];
      out[half * 8 + k] = pos;
    }
  }
  double det = s[0] * m[0][0] + s[1] * m[0][1] + s[2] * m[0][2] + s[3] * m[0][3];
  det = 1.0f / det;
  for (int k = 0; k < 16; k++) out[k] *= det;
}

// Matrix::Mult (matrix.cc:195-204): dst[i][j] = sum_k m0[k][j] * m1[i][k], accumulated from 0.
void multiply(Mat4 dst, Mat4 m0, Mat4 m1) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double acc = 0;
      for (int k = 0; k < 4; k++) acc += m0[k][j] * m1[i][k];
      dst[i][j] = acc;
    }
}

// Matrix::MultV (matrix.cc:206-216): row-vector transform with translation row 3.
void transform_point(double dst[3], Mat4 m, const double v[3]) {
  for (int c = 0; c < 3; c++) dst[c] = m[0][c] * v[0] + m[1][c] * v[1] + m[2][c] * v[2] + m[3][c];
}

// build_rotmatrix (trackball.cc:272-292)
void quat_matrix(Mat4 m, const double q[4]) {
  m[0][0] = 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]);
  m[0][1] = 2.0 * (q[0] * q[1] - q[2] * q[3]);
  m[0][2] = 2.0 * (q[2] * q[0] + q[1] * q[3]);
  m[1][0] = 2.0 * (q[0] * q[1] + q[2] * q[3]);
  m[1][1] = 1.0 - 2.0 * (q[2] * q[2] + q[0] * q[0]);
  m[1][2] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
  m[2][0] = 2.0 * (q[2] * q[0] - q[1] * q[3]);
  m[2][1] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
  m[2][2] = 1.0 - 2.0 * (q[1] * q[1] + q[0] * q[0]);
  m[0][3] = m[1][3] = m[2][3] = 0.0;
  m[3][0] = m[3][1] = m[3][2] = 0.0;
  m[3][3] = 1.0;
}

} // namespace

extern "C" int mb200_camera_frame_build(mb200_camera_frame *out, const double eye[3], const double lookat[3],
                                        const double up[3], double fov, const double quat[4], int width,
                                        int height) {
  if (!out || !eye || !lookat || !up || !quat || width <= 0 || height <= 0) return MB200_ERR_INVALID_ARG;
  Mat4 rot, local, m;
  quat_matrix(rot, quat);

  const double to_target[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
  const double dist = length3(to_target);
  double fwd[3] = {0.0, 0.0, dist};
  invert(rot);

  const double zero[3] = {0.0, 0.0, 0.0}, y_up[3] = {0.0, 1.0, 0.0};
  look_at(local, fwd, zero, y_up);
  local[3][0] += eye[0];
  local[3][1] += eye[1];
  local[3][2] += (eye[2] - dist);
  multiply(m, rot, local);

  double eye1[3], lookat1[3];
  transform_point(eye1, m, zero);
  fwd[2] = -fwd[2];
  transform_point(lookat1, m, fwd);

  // camera.cc:142-144: the transformed up vector is discarded in favour of the caller's.
  const double up1[3] = {up[0], up[1], up[2]};

  const double flen = (0.5f * (double)height / tanf(0.5f * (double)(fov * M_PI / 180.0f)));
  double look1[3] = {lookat1[0] - eye1[0], lookat1[1] - eye1[1], lookat1[2] - eye1[2]};
  double *u = out->du, *v = out->dv;
  cross3(u, look1, up1);
  normalize_f(u);
  cross3(v, look1, u);
  normalize_f(v);
  normalize_f(look1);
  for (int c = 0; c < 3; c++) look1[c] = flen * look1[c] + eye1[c];
  for (int c = 0; c < 3; c++) out->corner[c] = look1[c] - 0.5f * (width * u[c] + height * v[c]);
  for (int c = 0; c < 3; c++) out->origin[c] = eye1[c];
  return MB200_OK;
}

namespace mallie {

void Camera::BuildCameraFrame(double origin[3], double corner[3], double u[3], double v[3], double fov,
                              const double quat[4], int width, int height) {
  width_ = width;
  height_ = height;
  mb200_camera_frame f;
  mb200_camera_frame_build(&f, eye_, lookat_, up_, fov, quat, width, height);
  for (int c = 0; c < 3; c++) {
    origin[c] = origin_[c] = f.origin[c];
    corner[c] = corner_[c] = f.corner[c];
    u[c] = du_[c] = f.du[c];
    v[c] = dv_[c] = f.dv[c];
  }
  fov_ = fov;
}

// Camera::GenerateRay (camera.cc:222-240), host version for single rays; the batched
// version is the raygen kernel (mb200_generate_rays).
Ray Camera::GenerateRay(double u, double v) const {
  real3 dir;
  dir[0] = (corner_[0] + u * du_[0] + v * dv_[0]) - origin_[0];
  dir[1] = (corner_[1] + u * du_[1] + v * dv_[1]) - origin_[1];
  dir[2] = (corner_[2] + u * du_[2] + v * dv_[2]) - origin_[2];
  dir.normalize();
  Ray ray;
  ray.org = real3(origin_[0], origin_[1], origin_[2]);
  ray.dir = dir;
  return ray;
}

// Camera::GenerateEnvRay (camera.cc:242-257): host version for single rays (same libm as the reference).
Ray Camera::GenerateEnvRay(double u, double v) const {
  const double theta = M_PI * (v / height_);
  const double phi = 2.0 * M_PI * (u / width_);
  Ray ray;
  ray.org = real3(origin_[0], origin_[1], origin_[2]);
  ray.dir = real3(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi));
  return ray;
}

// Camera::GenerateStereoEnvRay (camera.cc:259-329): upper half of the image = left eye.
Ray Camera::GenerateStereoEnvRay(double u, double v) const {
  const bool left = v < (height_ >> 1);
  const double focal_length = 4.0, r = 0.5;
  const double theta = M_PI * fmod(2.0 * v / height_, 1.0);
  const double phi = 2.0 * M_PI * (u / width_);
  const real3 d0(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi));
  real3 parallax = left ? real3(-d0.z, 0.0, d0.x) : real3(d0.z, 0.0, -d0.x);
  parallax.normalize();
  parallax = parallax * r;
  Ray ray;
  ray.org = real3(origin_[0] + parallax.x, origin_[1] + parallax.y, origin_[2] + parallax.z);
  double psi = atan2(r, focal_length);
  if (left) psi = -psi;
  ray.dir = real3(d0.x * cos(psi) - d0.z * sin(psi), d0.y, d0.x * sin(psi) + d0.z * cos(psi));
  ray.dir.normalize();
  return ray;
}

} // namespace mallie

// oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" shim around the UNMODIFIED Mallie reference sources
// (compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libmallie_ref.so).  It exposes the reference's own
//   MeshLoader::LoadObj / LoadESON      (importers/mesh_loader.cc:26,212)
//   BVHAccel::Build / Traverse / Dump   (bvh_accel.cc:445,773,484)
//   Camera::BuildCameraFrame / GenerateRay (camera.cc:40,222)
//   Plane::intersect                    (prim-plane.cc:8)
//   mallie::Render                      (render.cc:593)
// to Python (ctypes) so that tests/ can pin oracle/mallie_oracle.c and the
// CUDA path against the real reference, and so that bench.py can time the
// reference's OpenMP CPU path (`--impl reference`, cpu_baseline.kind =
// "reference").  No reference source is copied: this file only #includes the
// reference headers through -I/root/reference.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <limits>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.h"
#include "mesh.h"
#include "intersection.h"
#include "bvh_accel.h"
#include "scene.h"
#include "camera.h"
#include "render.h"
#include "prim-plane.h"
#include "importers/mesh_loader.h"

namespace {

// Scene keeps mesh_/accel_ protected (scene.h:67-76); a subclass is the only
// unmodified way to hand it an in-memory mesh.
class HarnessScene : public mallie::Scene {
public:
  Mesh &mesh() { return mesh_; }
  BVHAccel &accel() { return accel_; }
};

struct RefScene {
  HarnessScene *scene;
};

static void zero_mesh(Mesh &m) { memset(&m, 0, sizeof(Mesh)); }

static double now_sec() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- sizes
int ref_sizeof_bvhnode(void) { return (int)sizeof(BVHNode); }
int ref_sizeof_intersection(void) { return (int)sizeof(Intersection); }
int ref_sizeof_ray(void) { return (int)sizeof(Ray); }
int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---------------------------------------------------------------- scene
// Create a scene from in-memory arrays (copied; Scene's destructor delete[]s
// vertices/faces/materialIDs, scene.cc:57-64).  normals/uvs nullable.
void *ref_scene_from_arrays(const double *vertices, size_t nverts,
                            const unsigned int *faces, size_t nfaces,
                            const unsigned int *materialIDs,
                            const double *fv_normals, const double *fv_uvs) {
  RefScene *s = new RefScene;
  s->scene = new HarnessScene;
  Mesh &m = s->scene->mesh();
  zero_mesh(m);
  m.numVertices = nverts;
  m.numFaces = nfaces;
  m.vertices = new real[3 * nverts + 1];
  memcpy(m.vertices, vertices, sizeof(real) * 3 * nverts);
  m.faces = new unsigned int[3 * nfaces + 1];
  memcpy(m.faces, faces, sizeof(unsigned int) * 3 * nfaces);
  if (materialIDs) {
    m.materialIDs = new unsigned int[nfaces + 1];
    memcpy(m.materialIDs, materialIDs, sizeof(unsigned int) * nfaces);
  }
  if (fv_normals) {
    m.facevarying_normals = new real[9 * nfaces + 1];
    memcpy(m.facevarying_normals, fv_normals, sizeof(real) * 9 * nfaces);
  }
  if (fv_uvs) {
    m.facevarying_uvs = new real[6 * nfaces + 1];
    memcpy(m.facevarying_uvs, fv_uvs, sizeof(real) * 6 * nfaces);
  }
  return s;
}

// kind: 0 = .obj (MeshLoader::LoadObj), 1 = .eson (MeshLoader::LoadESON).
// Applies scene_scale exactly as Scene::Init does (scene.cc:162-170) when
// scale != 1.0 (scene_fit not exposed here).
void *ref_scene_from_file(const char *path, int kind, double scene_scale) {
  RefScene *s = new RefScene;
  s->scene = new HarnessScene;
  Mesh &m = s->scene->mesh();
  zero_mesh(m);
  bool ok = (kind == 0) ? MeshLoader::LoadObj(m, path)
                        : MeshLoader::LoadESON(m, path);
  if (!ok) {
    zero_mesh(m);
    delete s->scene;
    delete s;
    return NULL;
  }
  for (size_t i = 0; i < m.numVertices; i++) {
    m.vertices[3 * i + 0] *= scene_scale;
    m.vertices[3 * i + 1] *= scene_scale;
    m.vertices[3 * i + 2] *= scene_scale;
  }
  return s;
}

// The real Scene::Init (scene.cc:66-250): load, scene_fit / scene_scale, BVHAccel::Build with default options.
// kind 0 = .obj, 1 = .eson.  bounds (nullable): Scene::BoundingBox after Init, bmin[3] then bmax[3].
void *ref_scene_init(const char *path, int kind, double scene_scale, int scene_fit, double *bounds) {
  RefScene *s = new RefScene;
  s->scene = new HarnessScene;
  zero_mesh(s->scene->mesh());
  const std::string p(path), none;
  const bool ok = s->scene->Init(kind == 0 ? p : none, kind == 1 ? p : none, none, none, scene_scale, scene_fit != 0);
  if (!ok) {
    zero_mesh(s->scene->mesh());
    delete s->scene;
    delete s;
    return NULL;
  }
  if (bounds) {
    real3 bmin, bmax;
    s->scene->BoundingBox(bmin, bmax);
    for (int k = 0; k < 3; k++) bounds[k] = bmin[k], bounds[3 + k] = bmax[k];
  }
  return s;
}

void ref_scene_destroy(void *h) {
  RefScene *s = (RefScene *)h;
  if (!s) return;
  // facevarying arrays are leaked by the reference's ~Scene; free them here.
  Mesh &m = s->scene->mesh();
  delete[] m.facevarying_normals;
  delete[] m.facevarying_uvs;
  m.facevarying_normals = NULL;
  m.facevarying_uvs = NULL;
  delete s->scene;
  delete s;
}

size_t ref_scene_num_vertices(void *h) { return ((RefScene *)h)->scene->mesh().numVertices; }
size_t ref_scene_num_faces(void *h) { return ((RefScene *)h)->scene->mesh().numFaces; }
int ref_scene_has_normals(void *h) { return ((RefScene *)h)->scene->mesh().facevarying_normals != NULL; }
int ref_scene_has_uvs(void *h) { return ((RefScene *)h)->scene->mesh().facevarying_uvs != NULL; }
int ref_scene_has_material_ids(void *h) { return ((RefScene *)h)->scene->mesh().materialIDs != NULL; }

void ref_scene_get_mesh(void *h, double *vertices, unsigned int *faces,
                        unsigned int *materialIDs, double *fv_normals,
                        double *fv_uvs) {
  Mesh &m = ((RefScene *)h)->scene->mesh();
  if (vertices) memcpy(vertices, m.vertices, sizeof(real) * 3 * m.numVertices);
  if (faces) memcpy(faces, m.faces, sizeof(unsigned int) * 3 * m.numFaces);
  if (materialIDs && m.materialIDs)
    memcpy(materialIDs, m.materialIDs, sizeof(unsigned int) * m.numFaces);
  if (fv_normals && m.facevarying_normals)
    memcpy(fv_normals, m.facevarying_normals, sizeof(real) * 9 * m.numFaces);
  if (fv_uvs && m.facevarying_uvs)
    memcpy(fv_uvs, m.facevarying_uvs, sizeof(real) * 6 * m.numFaces);
}

// BVHAccel::Build with default BVHBuildOptions (scene.cc:224-230). Returns
// build seconds, <0 on failure.
double ref_scene_build(void *h) {
  RefScene *s = (RefScene *)h;
  BVHBuildOptions options;
  double t0 = now_sec();
  bool ok = s->scene->accel().Build(&s->scene->mesh(), options);
  double t1 = now_sec();
  return ok ? (t1 - t0) : -1.0;
}

// The same with explicit BVHBuildOptions (bvh_accel.h:32-42).
double ref_scene_build_opts(void *h, double cost_taabb, int min_leaf_primitives, int max_tree_depth, int bin_size) {
  RefScene *s = (RefScene *)h;
  BVHBuildOptions options;
  options.costTaabb = cost_taabb;
  options.minLeafPrimitives = min_leaf_primitives;
  options.maxTreeDepth = max_tree_depth;
  options.binSize = bin_size;
  double t0 = now_sec();
  bool ok = s->scene->accel().Build(&s->scene->mesh(), options);
  double t1 = now_sec();
  return ok ? (t1 - t0) : -1.0;
}

size_t ref_scene_num_nodes(void *h) { return ((RefScene *)h)->scene->accel().GetNodes().size(); }
size_t ref_scene_num_indices(void *h) { return ((RefScene *)h)->scene->accel().GetIndices().size(); }

void ref_scene_get_bvh(void *h, void *nodes, unsigned int *indices) {
  BVHAccel &a = ((RefScene *)h)->scene->accel();
  const std::vector<BVHNode> &n = a.GetNodes();
  const std::vector<unsigned int> &ix = a.GetIndices();
  if (nodes && !n.empty()) memcpy(nodes, &n[0], sizeof(BVHNode) * n.size());
  if (indices && !ix.empty()) memcpy(indices, &ix[0], sizeof(unsigned int) * ix.size());
}

void ref_scene_get_stats(void *h, int *out3) {
  BVHBuildStatistics st = ((RefScene *)h)->scene->accel().GetStatistics();
  out3[0] = st.maxTreeDepth;
  out3[1] = st.numLeafNodes;
  out3[2] = st.numBranchNodes;
}

int ref_scene_dump(void *h, const char *path) {
  return ((RefScene *)h)->scene->accel().Dump(path) ? 1 : 0;
}
int ref_scene_load(void *h, const char *path) {
  return ((RefScene *)h)->scene->accel().Load(path) ? 1 : 0;
}

// ---------------------------------------------------------------- trace
// rays: n x {org[3], dir[3]} doubles.  hits: n x 32 B {t,u,v,faceID,matID}
// (head of Intersection, intersection.h:6-11).  isects (nullable): n full
// 184-byte Intersection records, zero-initialised before the call because
// the reference leaves untouched fields indeterminate.  hitmask (nullable):
// the bool Scene::Trace returned.  OpenMP schedule(dynamic,1) over chunks of
// `row` rays, mirroring render.cc:657.  Returns the best wall time over
// `repeat` runs.
double ref_scene_trace(void *h, const double *rays, size_t n, void *hits,
                       void *isects, unsigned char *hitmask, int row,
                       int nthreads, int repeat) {
  RefScene *s = (RefScene *)h;
  struct HitRec { double t, u, v; unsigned int faceID, materialID; };
  HitRec *out = (HitRec *)hits;
  Intersection *full = (Intersection *)isects;
  if (row <= 0) row = 1920;
  long nrows = (long)((n + row - 1) / row);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  double best = 1e300;
  for (int r = 0; r < (repeat > 0 ? repeat : 1); r++) {
    double t0 = now_sec();
#pragma omp parallel for schedule(dynamic, 1)
    for (long y = 0; y < nrows; y++) {
      size_t b = (size_t)y * row, e = b + row;
      if (e > n) e = n;
      for (size_t i = b; i < e; i++) {
        Ray ray;
        ray.org = real3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]);
        ray.dir = real3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        Intersection isect;
        memset(&isect, 0, sizeof(isect));
        bool hit = s->scene->Trace(isect, ray);
        if (out) {
          out[i].t = isect.t;
          out[i].u = isect.u;
          out[i].v = isect.v;
          out[i].faceID = isect.faceID;
          out[i].materialID = isect.materialID;
        }
        if (full) full[i] = isect;
        if (hitmask) hitmask[i] = hit ? 1 : 0;
      }
    }
    double t1 = now_sec();
    if (t1 - t0 < best) best = t1 - t0;
  }
  return best;
}

// ---------------------------------------------------------------- camera
void ref_camera_frame(const double eye[3], const double lookat[3],
                      const double up[3], double fov, const double quat[4],
                      int width, int height, double origin[3],
                      double corner[3], double du[3], double dv[3]) {
  mallie::Camera cam(eye, lookat, up);
  cam.BuildCameraFrame(origin, corner, du, dv, fov, quat, width, height);
}

// rays[i] = GenerateRay(px[i], py[i]) for a camera with the given frame.
void ref_camera_generate(const double eye[3], const double lookat[3],
                         const double up[3], double fov, const double quat[4],
                         int width, int height, const double *px,
                         const double *py, size_t n, double *rays) {
  mallie::Camera cam(eye, lookat, up);
  double origin[3], corner[3], du[3], dv[3];
  cam.BuildCameraFrame(origin, corner, du, dv, fov, quat, width, height);
  for (size_t i = 0; i < n; i++) {
    Ray r = cam.GenerateRay(px[i], py[i]);
    rays[6 * i + 0] = r.org[0];
    rays[6 * i + 1] = r.org[1];
    rays[6 * i + 2] = r.org[2];
    rays[6 * i + 3] = r.dir[0];
    rays[6 * i + 4] = r.dir[1];
    rays[6 * i + 5] = r.dir[2];
  }
}

// rays[i] = GenerateEnvRay / GenerateStereoEnvRay (px[i], py[i]) (camera.cc:242-329).
void ref_camera_generate_env(const double eye[3], const double lookat[3], const double up[3], double fov,
                             const double quat[4], int width, int height, const double *px, const double *py,
                             size_t n, int stereo, double *rays) {
  mallie::Camera cam(eye, lookat, up);
  double origin[3], corner[3], du[3], dv[3];
  cam.BuildCameraFrame(origin, corner, du, dv, fov, quat, width, height);
  for (size_t i = 0; i < n; i++) {
    Ray r = stereo ? cam.GenerateStereoEnvRay(px[i], py[i]) : cam.GenerateEnvRay(px[i], py[i]);
    for (int k = 0; k < 3; k++) rays[6 * i + k] = r.org[k], rays[6 * i + 3 + k] = r.dir[k];
  }
}

// Un-jittered primary rays for every integer pixel, row-major.
void ref_camera_generate_grid(const double eye[3], const double lookat[3],
                              const double up[3], double fov,
                              const double quat[4], int width, int height,
                              double *rays) {
  mallie::Camera cam(eye, lookat, up);
  double origin[3], corner[3], du[3], dv[3];
  cam.BuildCameraFrame(origin, corner, du, dv, fov, quat, width, height);
#pragma omp parallel for
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      Ray r = cam.GenerateRay((double)x, (double)y);
      size_t i = (size_t)y * width + x;
      rays[6 * i + 0] = r.org[0];
      rays[6 * i + 1] = r.org[1];
      rays[6 * i + 2] = r.org[2];
      rays[6 * i + 3] = r.dir[0];
      rays[6 * i + 4] = r.dir[1];
      rays[6 * i + 5] = r.dir[2];
    }
  }
}

// ---------------------------------------------------------------- plane
// Plane::intersect on a batch. t_in[i] is the incoming isect.t; outputs
// t_out, position, normal, hit flag (prim-plane.cc:8-44).
void ref_plane_intersect(float a, float b, float c, float d,
                         const double *rays, const double *t_in, size_t n,
                         double *t_out, double *position, double *normal,
                         unsigned char *hit) {
  mallie::Plane pl;
  pl.set(a, b, c, d);
  for (size_t i = 0; i < n; i++) {
    Ray ray;
    ray.org = real3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]);
    ray.dir = real3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
    Intersection isect;
    memset(&isect, 0, sizeof(isect));
    isect.t = t_in[i];
    bool h = pl.intersect(&isect, ray);
    hit[i] = h ? 1 : 0;
    t_out[i] = isect.t;
    for (int k = 0; k < 3; k++) {
      position[3 * i + k] = isect.position[k];
      normal[3 * i + k] = isect.normal[k];
    }
  }
}

// ---------------------------------------------------------------- render
// One call of mallie::Render (render.cc:593).  NOT re-entrant and keeps
// function-static state (initial_pass, gPlane): `plane` is latched by the
// first call in the process.  image: float[3*W*H], count: int[W*H] (in/out).
// Returns wall seconds of the call.
double ref_scene_render(void *h, int width, int height, double fov,
                        const double eye[3], const double lookat[3],
                        const double up[3], const double quat[4], int plane,
                        int step, int nthreads, float *image, int *count) {
  RefScene *s = (RefScene *)h;
  mallie::RenderConfig cfg;
  cfg.width = width;
  cfg.height = height;
  cfg.fov = fov;
  cfg.plane = plane != 0;
  for (int k = 0; k < 3; k++) {
    cfg.eye[k] = eye[k];
    cfg.lookat[k] = lookat[k];
    cfg.up[k] = up[k];
  }
  for (int k = 0; k < 4; k++) cfg.quat[k] = quat[k];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<float> img((size_t)3 * width * height);
  std::vector<int> cnt((size_t)width * height);
  memcpy(&cnt[0], count, sizeof(int) * cnt.size());
  double t0 = now_sec();
  mallie::Render(*s->scene, cfg, img, cnt, eye, lookat, up, quat, step);
  double t1 = now_sec();
  printf("\n");
  memcpy(image, &img[0], sizeof(float) * img.size());
  memcpy(count, &cnt[0], sizeof(int) * cnt.size());
  return t1 - t0;
}

// One call of mallie::RenderPanoramic (render.cc:710-763): 10 samples per pixel accumulated into image,
// count += 10.  Same static-state caveats as ref_scene_render.
double ref_scene_render_panoramic(void *h, int width, int height, double fov, const double eye[3],
                                  const double lookat[3], const double up[3], const double quat[4], int stereo,
                                  int nthreads, float *image, int *count) {
  RefScene *s = (RefScene *)h;
  mallie::RenderConfig cfg;
  cfg.width = width;
  cfg.height = height;
  cfg.fov = fov;
  for (int k = 0; k < 3; k++) cfg.eye[k] = eye[k], cfg.lookat[k] = lookat[k], cfg.up[k] = up[k];
  for (int k = 0; k < 4; k++) cfg.quat[k] = quat[k];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<float> img((size_t)3 * width * height);
  std::vector<int> cnt((size_t)width * height);
  memcpy(&cnt[0], count, sizeof(int) * cnt.size());
  double t0 = now_sec();
  mallie::RenderPanoramic(*s->scene, cfg, img, cnt, eye, lookat, up, quat, stereo != 0);
  double t1 = now_sec();
  printf("\n");
  memcpy(image, &img[0], sizeof(float) * img.size());
  memcpy(count, &cnt[0], sizeof(int) * cnt.size());
  return t1 - t0;
}

} // extern "C"

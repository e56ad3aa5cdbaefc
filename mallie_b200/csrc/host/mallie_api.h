// mallie_api.h -- host-side C++ mirror of the Mallie interfaces that sit on the
// render hot path, re-implemented over the mallie_b200 C ABI (include/mallie_b200.h).
//
// A program written against the reference's headers keeps compiling against
// these: same names, same argument meaning, same bool/assert error behaviour.
//   real, real3, vcross, vdot, Ray           <- common.h:6-83
//   Mesh                                     <- mesh.h:7-18
//   Intersection                             <- intersection.h:6-24
//   Material                                 <- material.h:6-24
//   BVHNode, BVHBuildOptions, BVHBuildStatistics, BVHAccel   <- bvh_accel.h:10-86
//   mallie::Camera                           <- camera.h:10-45
//   mallie::Scene                            <- scene.h:43-77
//   mallie::RenderConfig, mallie::Render     <- render.h:11-55
// The forwarding headers next to this file (common.h, mesh.h, scene.h, ...) let
// `#include "scene.h"` style code build unchanged.
//
// What differs from the reference (by design):
//   * BVHAccel::Traverse / Scene::Trace run on the GPU.  A single-ray call works
//     but pays a kernel launch; the batched TraceBatch()/Render() entries are
//     the intended use.
//   * Scene owns a device replica (mb200_scene) created at the end of Init().
#ifndef MALLIE_B200_HOST_API_H_
#define MALLIE_B200_HOST_API_H_

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

#include "mallie_b200.h"

// ----------------------------------------------------------------------------
// common.h
// ----------------------------------------------------------------------------
typedef double real;

struct real3 {
  real x, y, z;

  real3() {}
  real3(real a, real b, real c) : x(a), y(b), z(c) {}
  real3(real *p) : x(p[0]), y(p[1]), z(p[2]) {}

  real operator[](int i) const { return (&x)[i]; }
  real &operator[](int i) { return (&x)[i]; }

  real3 operator+(const real3 &o) const { return real3(x + o.x, y + o.y, z + o.z); }
  real3 operator-(const real3 &o) const { return real3(x - o.x, y - o.y, z - o.z); }
  real3 operator*(const real3 &o) const { return real3(x * o.x, y * o.y, z * o.z); }
  real3 operator/(const real3 &o) const { return real3(x / o.x, y / o.y, z / o.z); }
  real3 operator*(real s) const { return real3(x * s, y * s, z * s); }
  real3 &operator+=(const real3 &o) {
    x += o.x, y += o.y, z += o.z;
    return *this;
  }
  real3 neg() { return real3(-x, -y, -z); }
  real length() { return sqrt(x * x + y * y + z * z); }
  // Only vectors longer than 1e-6 are rescaled (common.h:48-57).
  void normalize() {
    real len = length();
    if (fabs(len) > 1.0e-6) {
      real inv = 1.0 / len;
      x *= inv, y *= inv, z *= inv;
    }
  }
};

inline real3 operator*(real s, const real3 &v) { return real3(v.x * s, v.y * s, v.z * s); }
inline real3 vcross(real3 a, real3 b) {
  return real3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline real vdot(real3 a, real3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct Ray {
  real3 org;
  real3 dir;
  real3 invDir;   // unused by Traverse (recomputed), kept for layout compatibility
  int dirSign[3]; // idem
};

// ----------------------------------------------------------------------------
// mesh.h -- field order and types as the reference (80-byte header).
// ----------------------------------------------------------------------------
typedef struct {
  size_t numVertices;
  size_t numFaces;
  real *vertices;                  // [xyz] * numVertices
  real *facevarying_normals;       // [xyz] * 3 * numFaces
  real *facevarying_tangents;      // unused on the path
  real *facevarying_binormals;     // unused on the path
  real *facevarying_uvs;           // [uv] * 3 * numFaces
  real *facevarying_vertex_colors; // unused on the path
  unsigned int *faces;             // 3 * numFaces
  unsigned int *materialIDs;       // numFaces
} Mesh;

// ----------------------------------------------------------------------------
// intersection.h -- 184 bytes; binary-compatible with mb200_isect.
// ----------------------------------------------------------------------------
typedef struct {
  real t, u, v;
  unsigned int faceID;
  unsigned int materialID;
  unsigned int f0, f1, f2;
  real3 position;
  real3 geometricNormal;
  real3 normal;
  real3 tangent;
  real3 binormal;
  real texcoord[2];
} Intersection;

static_assert(sizeof(Intersection) == sizeof(mb200_isect), "Intersection must match mb200_isect");
static_assert(sizeof(Ray) == 88, "Ray layout");

// ----------------------------------------------------------------------------
// material.h
// ----------------------------------------------------------------------------
struct Material {
  real3 diffuse, reflection, refraction;
  int id;
  Material() : diffuse(0.5, 0.5, 0.5), reflection(0.0, 0.0, 0.0), refraction(0.0, 0.0, 0.0), id(-1) {}
};

// ----------------------------------------------------------------------------
// bvh_accel.h
// ----------------------------------------------------------------------------
class BVHNode {
public:
  BVHNode() {}
  ~BVHNode() {}
  real bmin[3];
  real bmax[3];
  int flag; // 1 = leaf, 0 = branch
  int axis;
  unsigned int data[2]; // leaf: {ntris, first index}; branch: {child0, child1}
};
static_assert(sizeof(BVHNode) == sizeof(mb200_bvh_node), "BVHNode must match mb200_bvh_node");

struct BVHBuildOptions {
  bool debugPrint;
  real costTaabb;
  int minLeafPrimitives;
  int maxTreeDepth;
  int binSize;
  BVHBuildOptions() : debugPrint(false), costTaabb(0.2), minLeafPrimitives(16), maxTreeDepth(256), binSize(64) {}
};

struct BVHBuildStatistics {
  int maxTreeDepth;
  int numLeafNodes;
  int numBranchNodes;
  BVHBuildStatistics() : maxTreeDepth(0), numLeafNodes(0), numBranchNodes(0) {}
};

class BVHAccel {
public:
  BVHAccel();
  ~BVHAccel();

  // Binned-SAH build, bit-identical tree to bvh_accel.cc:445: on the GPU together with the scene upload
  // (mb200_scene_build) when one is present, else on the host (upload on first use).  SetDevice() first.
  bool Build(const Mesh *mesh, const BVHBuildOptions &options);
  BVHBuildStatistics GetStatistics() const { return stats_; }
  bool Dump(const char *filename);
  bool Load(const char *filename);

  // Closest hit for one ray (GPU launch of size 1).  Semantics of bvh_accel.cc:773-844.
  bool Traverse(Intersection &isect, const Mesh *mesh, Ray &ray);
  // Batched closest hit: isects[i] for rays[i]; returns number of hits, -1 on error.
  long TraverseBatch(Intersection *isects, const Mesh *mesh, const Ray *rays, size_t n, unsigned char *hitMask = 0);

  const std::vector<BVHNode> &GetNodes() const { return nodes_; }
  const std::vector<unsigned int> &GetIndices() const { return indices_; }

  // --- additions -----------------------------------------------------------
  void SetDevice(int device) {
    if (device != device_) ReleaseDevice(); // re-created from nodes_ / indices_ on the new GPU at first use
    device_ = device;
  }
  // Device replica; created lazily from (mesh, nodes_, indices_) on first use.
  mb200_scene *DeviceScene(const Mesh *mesh);
  // `count` replicas on GPUs device, device + 1, ... (element 0 is DeviceScene()); false if one cannot be made.
  bool DeviceScenes(const Mesh *mesh, int count, std::vector<mb200_scene *> &out);
  void ReleaseDevice();

private:
  BVHBuildOptions options_;
  std::vector<BVHNode> nodes_;
  std::vector<unsigned int> indices_;
  BVHBuildStatistics stats_;
  int device_;
  mb200_scene *dev_;
  const Mesh *devMesh_;
  std::vector<mb200_scene *> replicas_; // GPUs device_ + 1, ...
};

// ----------------------------------------------------------------------------
// mesh loading (importers/mesh_loader.h).  LoadObj follows the reference
// loader's vertex/face ordering so faceIDs agree.
// ----------------------------------------------------------------------------
class MeshLoader {
public:
  static bool LoadObj(Mesh &mesh, const char *filename);
  static bool LoadESON(Mesh &mesh, const char *filename);
};

namespace mallie {

// ----------------------------------------------------------------------------
// camera.h
// ----------------------------------------------------------------------------
class Camera {
public:
  Camera(const double eye[3], const double lookat[3], const double up[3]) {
    for (int i = 0; i < 3; i++) eye_[i] = eye[i], up_[i] = up[i], lookat_[i] = lookat[i];
  }
  ~Camera() {}

  void BuildCameraFrame(double origin[3], double corner[3], double u[3], double v[3], double fov,
                        const double quat[4], int width, int height);
  Ray GenerateRay(double u, double v) const;
  Ray GenerateEnvRay(double u, double v) const;       // camera.cc:242-257
  Ray GenerateStereoEnvRay(double u, double v) const; // camera.cc:259-329

  double eye_[3];
  double up_[3];
  double lookat_[3];
  double origin_[3];
  double corner_[3];
  double du_[3];
  double dv_[3];
  double fov_;
  int height_;
  int width_;
};

// ----------------------------------------------------------------------------
// scene.h
// ----------------------------------------------------------------------------
class Scene {
public:
  Scene();
  ~Scene();

  bool Init(const std::string &objFilename, const std::string &esonFilename,
            const std::string &magicaVoxelFilename, const std::string &materialFilename,
            double sceneScale = 1.0, bool sceneFit = false);
  // In-memory variant (arrays are copied): used by tools and tests.
  bool InitFromArrays(const double *vertices, size_t nverts, const unsigned int *faces, size_t nfaces,
                      const unsigned int *materialIDs, const double *fvNormals, const double *fvUVs);

  bool Trace(Intersection &isect, Ray &ray);
  long TraceBatch(Intersection *isects, const Ray *rays, size_t n, unsigned char *hitMask = 0);

  void BoundingBox(real3 &bmin, real3 &bmax);
  real3 GetBackgroundRadiance(real3 &dir);
  const Material &GetMaterial(int matID) const {
    static Material s_default;
    if (matID >= 0 && (size_t)matID < materials_.size()) return materials_[matID];
    return s_default;
  }

  // --- additions -----------------------------------------------------------
  void SetDevice(int device) { accel_.SetDevice(device); }
  mb200_scene *DeviceScene() { return accel_.DeviceScene(&mesh_); }
  bool DeviceScenes(int count, std::vector<mb200_scene *> &out) { return accel_.DeviceScenes(&mesh_, count, out); }
  const Mesh &GetMesh() const { return mesh_; }
  BVHAccel &GetAccel() { return accel_; }

protected:
  Mesh mesh_;
  std::vector<Material> materials_;
  BVHAccel accel_;
};

// ----------------------------------------------------------------------------
// render.h
// ----------------------------------------------------------------------------
struct RenderConfig {
  double fov;
  int width;
  int height;
  double eye[3];
  double lookat[3];
  double up[3];
  double quat[4];
  double scene_scale;
  bool scene_fit;
  bool plane;
  int num_passes;
  int num_photons;
  std::string obj_filename;
  std::string eson_filename;
  std::string magicavoxel_filename;
  std::string material_filename;

  // --- additions (all default to the reference's behaviour) ----------------
  int max_path_length; // kMaxPathLength (render.cc:52)
  int shader;          // mb200_shader; 0 = PathTrace
  double light[3];     // point light for the primary+shadow shader
  int device;          // CUDA device ordinal
  int num_gpus;        // >1: image rows are split across GPUs (single process)

  RenderConfig()
      : fov(45.0), width(512), height(512), scene_scale(1.0), scene_fit(false), plane(false), num_passes(10),
        num_photons(10000), max_path_length(16), shader(0), device(0), num_gpus(1) {
    eye[0] = 0.0, eye[1] = 0.0, eye[2] = -5.0;
    lookat[0] = lookat[1] = lookat[2] = 0.0;
    up[0] = 0.0, up[1] = 1.0, up[2] = 0.0;
    quat[0] = quat[1] = quat[2] = quat[3] = 0.0;
    light[0] = 0.0, light[1] = 20.0, light[2] = 0.0;
  }
};

// One pass: image (RGB float, >= 3*W*H) is zeroed then overwritten, count[p]++ when step == 1
// (render.cc:593-708).  Blocking; prints the "[Mallie] Render time" line.
void Render(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
            const double eye[3], const double lookat[3], const double up[3], const double quat[4], int step);

// RenderPanoramic (render.h:56-61, render.cc:710-763): image zeroed, then TEN samples of PathTraceEnv per pixel
// through the equirectangular (or top/bottom stereo) panorama camera are added, count += 10.
void RenderPanoramic(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
                     const double eye[3], const double lookat[3], const double up[3], const double quat[4],
                     bool stereo);

// num_passes passes accumulated on the GPU (the SDL render thread's loop, main_sdl.cc:572-606):
// image += sum of passes, count += num_passes.  Returns Mrays/s of the call.
double RenderAccumulate(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
                        const double eye[3], const double lookat[3], const double up[3], const double quat[4],
                        int num_passes, mb200_render_stats *stats = 0);
// Render + HDRToLDR of DoMainConsole (main_console.cc:57-75) in one device-side step (mb200_render_frame_ldr): only
// the 8-bit image is copied to the host.  ldr_mode: MB200_LDR_RGB8_LINEAR (HDRToLDR) or MB200_LDR_BGRA8_GAMMA22 (Display).
double RenderLDR(Scene &scene, const RenderConfig &config, std::vector<unsigned char> &out, const double eye[3],
                 const double lookat[3], const double up[3], const double quat[4], int num_passes, int ldr_mode,
                 mb200_render_stats *stats = 0);

// config.json -> RenderConfig (main.cc:98-205): same keys, unknown keys ignored.
bool LoadJSONConfig(RenderConfig &config, const std::string &filename);
bool LoadJSONConfigFromString(RenderConfig &config, const std::string &json_text);

// DoMainConsole (main_console.cc:57-75): one Render() pass -> 8-bit image -> file.  The reference
// writes output.jpg through jpge; this writes a binary PPM with the same quantisation
// (x / count * 255.5, clamped, main_console.cc:25-43).  `passes` > 1 accumulates that many passes on
// the GPU first (what the SDL front end's render thread does).  Returns false on failure.
bool DoMainConsole(Scene &scene, const RenderConfig &config, const char *output = "output.ppm", int passes = 1);
void HDRToLDR(std::vector<unsigned char> &out, const std::vector<float> &in, const std::vector<int> &in_count,
              int width, int height);

} // namespace mallie

#endif // MALLIE_B200_HOST_API_H_

// mallie_api.cc -- the host C++ classes a Mallie program uses around the render hot path
// (BVHAccel, Scene, Render), implemented over the mallie_b200 C ABI (include/mallie_b200.h).
// Interfaces, argument meaning and bool/printf error behaviour follow the reference:
//   BVHAccel   bvh_accel.h:54-86, bvh_accel.cc:445-544,773-844
//   Scene      scene.h:43-77,     scene.cc:52-333
//   Render     render.h:51-55,    render.cc:593-708
// There is no CPU tracing path in here: every Trace/Traverse/Render call runs on the GPU, and fails
// (false / error message) when no device scene can be created.
#include "mallie_api.h"

#include <cstdlib>

#include <cassert>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "mesh_data.h"

// ----------------------------------------------------------------------------------------------------
// BVHAccel
// ----------------------------------------------------------------------------------------------------
BVHAccel::BVHAccel() : device_(0), dev_(nullptr), devMesh_(nullptr) {}

BVHAccel::~BVHAccel() { ReleaseDevice(); }

void BVHAccel::ReleaseDevice() {
  for (mb200_scene *r : replicas_) mb200_scene_destroy(r);
  replicas_.clear();
  if (dev_) mb200_scene_destroy(dev_);
  dev_ = nullptr;
  devMesh_ = nullptr;
}

bool BVHAccel::DeviceScenes(const Mesh *mesh, int count, std::vector<mb200_scene *> &out) {
  out.clear();
  mb200_scene *first = DeviceScene(mesh);
  if (!first || count < 1) return false;
  out.push_back(first);
  while ((int)replicas_.size() < count - 1) {
    mb200_scene *r = nullptr;
    const int rc = mb200_scene_clone(&r, first, device_ + 1 + (int)replicas_.size()); // device-to-device copy
    if (rc != MB200_OK) {
      printf("Mallie:err\tmsg:cannot create the scene replica on GPU %d: %s\n", device_ + 1 + (int)replicas_.size(),
             mb200_last_error());
      return false;
    }
    replicas_.push_back(r);
  }
  for (int g = 1; g < count; g++) out.push_back(replicas_[g - 1]);
  return true;
}

// BVHAccel::Build (bvh_accel.cc:445-482).  With a GPU present the tree is grown on the device and the traversal
// layout is written from it there (mb200_scene_build; the reference-layout tree is downloaded into nodes_ /
// indices_ for GetNodes / Dump); without one -- or with MB200_HOST_BUILD=1, or minLeafPrimitives < 2 -- the host
// builder runs and the upload happens on first use.  Either way the tree is bit-identical to the reference's.
bool BVHAccel::Build(const Mesh *mesh, const BVHBuildOptions &options) {
  assert(mesh);
  options_ = options;
  ReleaseDevice();
  mb200_build_options o;
  o.cost_taabb = options.costTaabb;
  o.min_leaf_primitives = options.minLeafPrimitives;
  o.max_tree_depth = options.maxTreeDepth;
  o.bin_size = options.binSize;
  mb200_bvh *b = nullptr;
  const char *force_host = getenv("MB200_HOST_BUILD");
  const bool on_device = !(force_host && atoi(force_host) != 0) && o.min_leaf_primitives >= 2 && mesh->numFaces > 0 &&
                         mb200_device_count() > device_;
  bool built = false;
  if (on_device) {
    const int rc = mb200_scene_build(&dev_, device_, mesh->vertices, mesh->numVertices, mesh->faces, mesh->numFaces,
                                     mesh->materialIDs, mesh->facevarying_normals, mesh->facevarying_uvs, &o, &b);
    if (rc == MB200_OK) {
      devMesh_ = mesh;
      built = true;
    } else {
      dev_ = nullptr;
      // a visible GPU that is not sm_100 class cannot build (or trace); the tree itself does not need one
      if (rc != MB200_ERR_NO_DEVICE) {
        printf("Mallie:err\tmsg:BVH build failed: %s\n", mb200_last_error());
        return false;
      }
    }
  }
  if (!built && mb200_bvh_build(&b, mesh->vertices, mesh->numVertices, mesh->faces, mesh->numFaces, &o) != MB200_OK) {
    printf("Mallie:err\tmsg:BVH build failed: %s\n", mb200_last_error());
    return false;
  }
  const size_t nn = mb200_bvh_num_nodes(b), ni = mb200_bvh_num_indices(b);
  nodes_.resize(nn);
  indices_.resize(ni);
  if (nn) memcpy(static_cast<void *>(nodes_.data()), mb200_bvh_nodes(b), nn * sizeof(BVHNode));
  if (ni) memcpy(indices_.data(), mb200_bvh_indices(b), ni * sizeof(unsigned int));
  mb200_build_stats st;
  mb200_bvh_stats(b, &st);
  stats_.maxTreeDepth = st.max_tree_depth;
  stats_.numLeafNodes = st.num_leaf_nodes;
  stats_.numBranchNodes = st.num_branch_nodes;
  mb200_bvh_destroy(b);
  return true;
}

// BVHAccel::Dump / Load (bvh_accel.cc:484-544): u64 numNodes, BVHNode[numNodes], u64 numIndices,
// u32[numIndices] -- byte-compatible with the reference's files.
bool BVHAccel::Dump(const char *filename) {
  FILE *fp = fopen(filename, "wb");
  if (!fp) {
    fprintf(stderr, "[BVHAccel] Cannot write a file: %s\n", filename);
    return false;
  }
  const unsigned long long nn = nodes_.size(), ni = indices_.size();
  bool ok = fwrite(&nn, sizeof(nn), 1, fp) == 1;
  ok = ok && (nn == 0 || fwrite(nodes_.data(), sizeof(BVHNode), nn, fp) == nn);
  ok = ok && fwrite(&ni, sizeof(ni), 1, fp) == 1;
  ok = ok && (ni == 0 || fwrite(indices_.data(), sizeof(unsigned int), ni, fp) == ni);
  fclose(fp);
  return ok;
}

bool BVHAccel::Load(const char *filename) {
  mb200_bvh *b = nullptr;
  if (mb200_bvh_load(&b, filename) != MB200_OK) {
    fprintf(stderr, "Cannot open file: %s\n", filename);
    return false;
  }
  ReleaseDevice();
  const size_t nn = mb200_bvh_num_nodes(b), ni = mb200_bvh_num_indices(b);
  nodes_.resize(nn);
  indices_.resize(ni);
  if (nn) memcpy(static_cast<void *>(nodes_.data()), mb200_bvh_nodes(b), nn * sizeof(BVHNode));
  if (ni) memcpy(indices_.data(), mb200_bvh_indices(b), ni * sizeof(unsigned int));
  mb200_build_stats st;
  mb200_bvh_stats(b, &st);
  stats_.maxTreeDepth = st.max_tree_depth;
  stats_.numLeafNodes = st.num_leaf_nodes;
  stats_.numBranchNodes = st.num_branch_nodes;
  mb200_bvh_destroy(b);
  return true;
}

mb200_scene *BVHAccel::DeviceScene(const Mesh *mesh) {
  if (dev_ && devMesh_ == mesh) return dev_;
  ReleaseDevice();
  if (!mesh) return nullptr;
  const int rc = mb200_scene_create(&dev_, device_, mesh->vertices, mesh->numVertices, mesh->faces, mesh->numFaces,
                                    mesh->materialIDs, mesh->facevarying_normals, mesh->facevarying_uvs,
                                    reinterpret_cast<const mb200_bvh_node *>(nodes_.data()), nodes_.size(),
                                    indices_.data(), indices_.size());
  if (rc != MB200_OK) {
    printf("Mallie:err\tmsg:cannot create the device scene: %s\n", mb200_last_error());
    dev_ = nullptr;
    return nullptr;
  }
  devMesh_ = mesh;
  return dev_;
}

// Batched BVHAccel::Traverse.  isects[i] is written as the reference writes it: on a hit every field
// BuildIntersection sets; on a miss only t = DBL_MAX, u = v = 0, faceID = -1 (bvh_accel.cc:783-786),
// the rest of the record is left as the caller had it.
long BVHAccel::TraverseBatch(Intersection *isects, const Mesh *mesh, const Ray *rays, size_t n,
                             unsigned char *hitMask) {
  if (n == 0) return 0;
  mb200_scene *s = DeviceScene(mesh);
  if (!s || !isects || !rays) return -1;
  std::vector<mb200_ray> packed(n);
  for (size_t i = 0; i < n; i++)
    for (int c = 0; c < 3; c++) packed[i].org[c] = rays[i].org[c], packed[i].dir[c] = rays[i].dir[c];
  std::vector<mb200_isect> out(n);
  std::vector<unsigned char> mask(n);
  if (mb200_trace_closest_full(s, packed.data(), n, out.data(), mask.data()) != MB200_OK) {
    printf("Mallie:err\tmsg:trace failed: %s\n", mb200_last_error());
    return -1;
  }
  long hits = 0;
  for (size_t i = 0; i < n; i++) {
    Intersection &d = isects[i];
    const mb200_isect &o = out[i];
    d.t = o.t, d.u = o.u, d.v = o.v, d.faceID = o.faceID;
    if (mask[i]) {
      hits++;
      d.materialID = o.materialID;
      d.f0 = o.f0, d.f1 = o.f1, d.f2 = o.f2;
      for (int c = 0; c < 3; c++) {
        d.position[c] = o.position[c];
        d.geometricNormal[c] = o.geometricNormal[c];
        d.normal[c] = o.normal[c];
      }
      d.texcoord[0] = o.texcoord[0], d.texcoord[1] = o.texcoord[1];
    }
    if (hitMask) hitMask[i] = mask[i];
  }
  return hits;
}

bool BVHAccel::Traverse(Intersection &isect, const Mesh *mesh, Ray &ray) {
  return TraverseBatch(&isect, mesh, &ray, 1, nullptr) == 1;
}

namespace mb200 {

void apply_scene_transform(double *v, size_t nverts, double scene_scale, bool scene_fit, bool verbose) {
  if (scene_fit) { // to [-1, 1]^3, scene.cc:112-160
    double bmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, bmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (size_t i = 0; i < nverts; i++)
      for (int c = 0; c < 3; c++) {
        bmin[c] = v[3 * i + c] < bmin[c] ? v[3 * i + c] : bmin[c];
        bmax[c] = bmax[c] < v[3 * i + c] ? v[3 * i + c] : bmax[c];
      }
    double inv[3];
    for (int c = 0; c < 3; c++) {
      const double ext = bmax[c] - bmin[c];
      inv[c] = (ext > 0.000001) ? (1.0 / ext) : ext;
    }
    if (verbose) {
      printf("bmin = %f, %f, %f\n", bmin[0], bmin[1], bmin[2]);
      printf("bmax = %f, %f, %f\n", bmax[0], bmax[1], bmax[2]);
      printf("binv = %f, %f, %f\n", inv[0], inv[1], inv[2]);
    }
    for (size_t i = 0; i < nverts; i++)
      for (int c = 0; c < 3; c++) {
        double x = v[3 * i + c];
        x -= bmin[c];
        x *= inv[c];
        x -= 0.5;
        x *= 2.0;
        v[3 * i + c] = x;
      }
  } else {
    for (size_t i = 0; i < 3 * nverts; i++) v[i] *= scene_scale;
  }
}

} // namespace mb200

namespace mallie {

// ----------------------------------------------------------------------------------------------------
// Scene
// ----------------------------------------------------------------------------------------------------
Scene::Scene() { memset(&mesh_, 0, sizeof(mesh_)); }

Scene::~Scene() {
  accel_.ReleaseDevice();
  delete[] mesh_.vertices;
  delete[] mesh_.faces;
  delete[] mesh_.materialIDs;
  delete[] mesh_.facevarying_normals;
  delete[] mesh_.facevarying_uvs;
}

static bool finish_init(Scene &scene, Mesh &mesh, BVHAccel &accel) {
  BVHBuildOptions options; // defaults, scene.cc:224
  printf("  BVH build option:\n");
  printf("    # of leaf primitives: %d\n", options.minLeafPrimitives);
  printf("    SAH binsize         : %d\n", options.binSize);
  const auto t0 = std::chrono::steady_clock::now();
  if (!accel.Build(&mesh, options)) return false;
  const BVHBuildStatistics stats = accel.GetStatistics();
  printf("  BVH statistics:\n");
  printf("    # of leaf   nodes: %d\n", stats.numLeafNodes);
  printf("    # of branch nodes: %d\n", stats.numBranchNodes);
  printf("  Max tree depth   : %d\n", stats.maxTreeDepth);
  const auto t1 = std::chrono::steady_clock::now();
  printf("  BVH build time: %d msecs\n", (int)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
  if (!accel.GetNodes().empty()) {
    real3 bmin, bmax;
    scene.BoundingBox(bmin, bmax);
    printf("  BVH bounding box:\n");
    printf("    bmin = (%f, %f, %f)\n", bmin[0], bmin[1], bmin[2]);
    printf("    bmax = (%f, %f, %f)\n", bmax[0], bmax[1], bmax[2]);
  }
  return true;
}

bool Scene::Init(const std::string &objFilename, const std::string &esonFilename,
                 const std::string &magicaVoxelFilename, const std::string &materialFilename, double sceneScale,
                 bool sceneFit) {
  (void)materialFilename; // parsed from config.json but never read by the reference either (scene.cc:66-251)
  bool ret = false;
  if (!objFilename.empty()) {
    ret = MeshLoader::LoadObj(mesh_, objFilename.c_str());
    if (!ret) {
      printf("Mallie:err\tmsg:Failed to load .obj file [ %s ]\n", objFilename.c_str());
      return false;
    }
    printf("Mallie:info\tmsg:Success to load .obj file [ %s ]\n", objFilename.c_str());
  } else if (!esonFilename.empty()) {
    ret = MeshLoader::LoadESON(mesh_, esonFilename.c_str());
    if (!ret) {
      printf("Mallie:err\tmsg:Failed to load .eson file [ %s ]\n", esonFilename.c_str());
      return false;
    }
    printf("Mallie:info\tmsg:Success to load .eson file [ %s ]\n", esonFilename.c_str());
  } else if (!magicaVoxelFilename.empty()) {
    // asset format outside the render hot path: not provided
    printf("Mallie:err\tmsg:Failed to load .vox file [ %s ]\n", magicaVoxelFilename.c_str());
    return false;
  }
  if (!ret) {
    printf("Mallie:err\tmsg:Failed to load mesh\n");
    return false;
  }
  mb200::apply_scene_transform(mesh_.vertices, mesh_.numVertices, sceneScale, sceneFit, true);
  return finish_init(*this, mesh_, accel_);
}

bool Scene::InitFromArrays(const double *vertices, size_t nverts, const unsigned int *faces, size_t nfaces,
                           const unsigned int *materialIDs, const double *fvNormals, const double *fvUVs) {
  if ((!vertices && nverts) || (!faces && nfaces)) return false;
  mesh_.numVertices = nverts;
  mesh_.numFaces = nfaces;
  mesh_.vertices = new real[3 * nverts + 1];
  mesh_.faces = new unsigned int[3 * nfaces + 1];
  mesh_.materialIDs = new unsigned int[nfaces + 1];
  if (nverts) memcpy(mesh_.vertices, vertices, 3 * nverts * sizeof(real));
  if (nfaces) memcpy(mesh_.faces, faces, 3 * nfaces * sizeof(unsigned int));
  for (size_t i = 0; i < nfaces; i++) mesh_.materialIDs[i] = materialIDs ? materialIDs[i] : 0u;
  if (fvNormals) {
    mesh_.facevarying_normals = new real[9 * nfaces + 1];
    memcpy(mesh_.facevarying_normals, fvNormals, 9 * nfaces * sizeof(real));
  }
  if (fvUVs) {
    mesh_.facevarying_uvs = new real[6 * nfaces + 1];
    memcpy(mesh_.facevarying_uvs, fvUVs, 6 * nfaces * sizeof(real));
  }
  return finish_init(*this, mesh_, accel_);
}

bool Scene::Trace(Intersection &isect, Ray &ray) { return accel_.Traverse(isect, &mesh_, ray); }

long Scene::TraceBatch(Intersection *isects, const Ray *rays, size_t n, unsigned char *hitMask) {
  return accel_.TraverseBatch(isects, &mesh_, rays, n, hitMask);
}

// Scene::BoundingBox (scene.cc:317-333): root node bounds.
void Scene::BoundingBox(real3 &bmin, real3 &bmax) {
  const std::vector<BVHNode> &nodes = accel_.GetNodes();
  assert(nodes.size() > 0);
  for (int c = 0; c < 3; c++) bmin[c] = nodes[0].bmin[c], bmax[c] = nodes[0].bmax[c];
}

real3 Scene::GetBackgroundRadiance(real3 &dir) {
  (void)dir;
  return real3(0.0, 0.0, 0.0);
}

// ----------------------------------------------------------------------------------------------------
// Render
// ----------------------------------------------------------------------------------------------------
namespace {

// The reference keeps the plane decision and the RNG in function-static / global state that is set
// up on the first call only (render.cc:113-116,615-628); Render is not re-entrant there either.
struct RenderState {
  bool initial_pass = true;
  bool plane = false;
  float plane_abcd[4] = {0, 0, 0, 0};
  unsigned int pass = 0; // replaces the per-thread RNG state: one stream per (pixel, pass)
};
RenderState g_render;

bool fill_params(mb200_render_params &p, Scene &scene, const RenderConfig &config, const double eye[3],
                 const double lookat[3], const double up[3], const double quat[4]) {
  mb200_render_params_default(&p, config.width, config.height);
  Camera camera(eye, lookat, up);
  double origin[3], corner[3], du[3], dv[3];
  camera.BuildCameraFrame(origin, corner, du, dv, config.fov, quat, config.width, config.height);
  for (int c = 0; c < 3; c++)
    p.frame.origin[c] = origin[c], p.frame.corner[c] = corner[c], p.frame.du[c] = du[c], p.frame.dv[c] = dv[c];
  if (g_render.initial_pass) {
    g_render.initial_pass = false;
    g_render.plane = config.plane;
    if (g_render.plane) {
      real3 bmin, bmax;
      scene.BoundingBox(bmin, bmax);
      const double lo[3] = {bmin[0], bmin[1], bmin[2]}, hi[3] = {bmax[0], bmax[1], bmax[2]};
      mb200_plane_from_bounds(lo, hi, g_render.plane_abcd);
    }
  }
  p.use_plane = g_render.plane ? 1 : 0;
  for (int k = 0; k < 4; k++) p.plane[k] = g_render.plane_abcd[k];
  p.max_path_length = config.max_path_length;
  p.shader = config.shader;
  for (int c = 0; c < 3; c++) p.light[c] = config.light[c];
  p.jitter = 1;
  return true;
}

} // namespace

void Render(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
            const double eye[3], const double lookat[3], const double up[3], const double quat[4], int step) {
  const int width = config.width, height = config.height;
  assert(image.size() >= (size_t)3 * width * height);
  assert(count.size() >= (size_t)width * height);
  const auto t0 = std::chrono::steady_clock::now();
  mb200_scene *s = scene.DeviceScene();
  if (!s) {
    printf("Mallie:err\tmsg:Render: no device scene (%s)\n", mb200_last_error());
    return;
  }
  mb200_render_params p;
  fill_params(p, scene, config, eye, lookat, up, quat);
  p.pass = g_render.pass++;
  p.pixel_step = step < 1 ? 1 : step;
  memset(image.data(), 0, sizeof(float) * (size_t)width * height * 3); // render.cc:639
  std::vector<mb200_scene *> gpus;
  if (config.num_gpus > 1 && p.pixel_step == 1 && scene.DeviceScenes(config.num_gpus, gpus)) {
    // rows interleaved over the GPUs in bands of 4 scanlines (one tile row); the frame is assembled on the first GPU
    std::vector<int> one((size_t)width * height);
    if (mb200_render_frame_multi(gpus.data(), (int)gpus.size(), &p, 1, 4, image.data(), one.data(), nullptr) != MB200_OK)
      printf("Mallie:err\tmsg:Render failed: %s\n", mb200_last_error());
    else
      for (size_t i = 0; i < one.size(); i++) count[i] += one[i];
  } else if (mb200_render_pass(s, &p, image.data(), count.data(), nullptr) != MB200_OK) {
    printf("Mallie:err\tmsg:Render failed: %s\n", mb200_last_error());
  }
  const auto t1 = std::chrono::steady_clock::now();
  const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  printf("\r[Mallie] Render time: %f sec(s) | %f fps", ms / 1000.0, 1000.0 / ms);
  fflush(stdout);
}

void RenderPanoramic(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
                     const double eye[3], const double lookat[3], const double up[3], const double quat[4],
                     bool stereo) {
  const int width = config.width, height = config.height;
  assert(image.size() >= (size_t)3 * width * height);
  assert(count.size() >= (size_t)width * height);
  const auto t0 = std::chrono::steady_clock::now();
  mb200_scene *s = scene.DeviceScene();
  if (!s) {
    printf("Mallie:err\tmsg:RenderPanoramic: no device scene (%s)\n", mb200_last_error());
    return;
  }
  mb200_render_params p;
  mb200_render_params_default(&p, width, height);
  Camera camera(eye, lookat, up);
  double origin[3], corner[3], du[3], dv[3];
  camera.BuildCameraFrame(origin, corner, du, dv, config.fov, quat, width, height);
  for (int c = 0; c < 3; c++)
    p.frame.origin[c] = origin[c], p.frame.corner[c] = corner[c], p.frame.du[c] = du[c], p.frame.dv[c] = dv[c];
  p.max_path_length = config.max_path_length;
  p.shader = MB200_SHADER_PATHTRACE_ENV;
  p.camera_mode = stereo ? MB200_CAMERA_ENV_STEREO : MB200_CAMERA_ENV;
  p.pass = g_render.pass;
  g_render.pass += 10u;
  memset(image.data(), 0, sizeof(float) * (size_t)width * height * 3); // render.cc:740
  if (mb200_render_accumulate(s, &p, 10, image.data(), count.data(), nullptr) != MB200_OK)
    printf("Mallie:err\tmsg:RenderPanoramic failed: %s\n", mb200_last_error());
  const auto t1 = std::chrono::steady_clock::now();
  const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  printf("\r[Mallie] Render time: %f sec(s) | %f fps", ms / 1000.0, 1000.0 / ms);
  fflush(stdout);
}

// DoMainConsole's Render + HDRToLDR (main_console.cc:57-75) in one device-side step: num_passes samples per pixel,
// quantised on the GPU; only the 8-bit image crosses PCIe.  Returns Mrays/s (0 on failure).
double RenderLDR(Scene &scene, const RenderConfig &config, std::vector<unsigned char> &out, const double eye[3],
                 const double lookat[3], const double up[3], const double quat[4], int num_passes, int ldr_mode,
                 mb200_render_stats *stats) {
  const int width = config.width, height = config.height;
  if (num_passes < 1 || width <= 0 || height <= 0) return 0.0;
  mb200_scene *s = scene.DeviceScene();
  if (!s) {
    printf("Mallie:err\tmsg:RenderLDR: no device scene (%s)\n", mb200_last_error());
    return 0.0;
  }
  out.resize((size_t)width * height * (ldr_mode == MB200_LDR_RGB8_LINEAR ? 3 : 4));
  mb200_render_params p;
  fill_params(p, scene, config, eye, lookat, up, quat);
  p.pass = g_render.pass;
  g_render.pass += (unsigned int)num_passes;
  mb200_render_stats local;
  const auto t0 = std::chrono::steady_clock::now();
  if (mb200_render_frame_ldr(s, &p, num_passes, ldr_mode, out.data(), &local) != MB200_OK) {
    printf("Mallie:err\tmsg:RenderLDR failed: %s\n", mb200_last_error());
    return 0.0;
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (stats) *stats = local;
  const double sec = std::chrono::duration<double>(t1 - t0).count();
  const double rays = (double)(local.primary_rays + local.bounce_rays + local.shadow_rays);
  return sec > 0.0 ? rays / sec / 1.0e6 : 0.0;
}

double RenderAccumulate(Scene &scene, const RenderConfig &config, std::vector<float> &image, std::vector<int> &count,
                        const double eye[3], const double lookat[3], const double up[3], const double quat[4],
                        int num_passes, mb200_render_stats *stats) {
  const int width = config.width, height = config.height;
  assert(image.size() >= (size_t)3 * width * height);
  assert(count.size() >= (size_t)width * height);
  if (num_passes < 1) return 0.0;
  mb200_scene *s = scene.DeviceScene();
  if (!s) {
    printf("Mallie:err\tmsg:RenderAccumulate: no device scene (%s)\n", mb200_last_error());
    return 0.0;
  }
  mb200_render_params p;
  fill_params(p, scene, config, eye, lookat, up, quat);
  p.pass = g_render.pass;
  g_render.pass += (unsigned int)num_passes;
  mb200_render_stats local;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<mb200_scene *> gpus;
  if (config.num_gpus > 1 && scene.DeviceScenes(config.num_gpus, gpus)) {
    std::vector<float> sum((size_t)width * height * 3);
    std::vector<int> cnt((size_t)width * height);
    if (mb200_render_frame_multi(gpus.data(), (int)gpus.size(), &p, num_passes, 4, sum.data(), cnt.data(), &local) != MB200_OK) {
      printf("Mallie:err\tmsg:RenderAccumulate failed: %s\n", mb200_last_error());
      return 0.0;
    }
    for (size_t i = 0; i < sum.size(); i++) image[i] += sum[i];
    for (size_t i = 0; i < cnt.size(); i++) count[i] += cnt[i];
  } else if (mb200_render_accumulate(s, &p, num_passes, image.data(), count.data(), &local) != MB200_OK) {
    printf("Mallie:err\tmsg:RenderAccumulate failed: %s\n", mb200_last_error());
    return 0.0;
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (stats) *stats = local;
  const double sec = std::chrono::duration<double>(t1 - t0).count();
  const double rays = (double)(local.primary_rays + local.bounce_rays + local.shadow_rays);
  return sec > 0.0 ? rays / sec / 1.0e6 : 0.0;
}

} // namespace mallie

// capi.cc -- the extern "C" boundary declared in include/mallie_b200.h.
//
// Thin glue only: argument checking, host<->device staging for callers that pass
// host buffers, stream ordering, error strings.  No algorithm lives here and there
// is NO CPU fallback: without a GPU every compute entry returns MB200_ERR_NO_DEVICE.
#include <cuda_runtime_api.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "device/kernels.h"
#include "device/scene.h"
#include "host/bvh_build.h"
#include "host/mallie_api.h"
#include "host/mesh_data.h"
#include "mallie_b200.h"

namespace {

thread_local std::string g_err;

int set_err(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

int cuda_err(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA;
}

#define CU(call)                                             \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return cuda_err(e_, #call);       \
  } while (0)

enum HostKind { kDevice = 0, kPinned = 1, kPageable = 2 };

// Where does a caller's buffer live?  Device/managed memory is used in place, page-locked host
// memory is DMA'd directly, pageable host memory goes through the scene's pinned staging.
HostKind classify(const void *p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return kPageable;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return kDevice;
  if (a.type == cudaMemoryTypeHost) return kPinned;
  return kPageable;
}

bool is_device_ptr(const void *p) { return p && classify(p) == kDevice; }

int ensure(mb200_scene::Staging &st, size_t bytes, bool need_pinned = true) {
  if (st.cap >= bytes && (!need_pinned || st.pinned)) return MB200_OK;
  if (st.pinned) cudaFreeHost(st.pinned);
  if (st.dev) cudaFree(st.dev);
  st.pinned = st.dev = nullptr;
  st.cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  CU(cudaMallocHost(&st.pinned, cap));
  CU(cudaMalloc(&st.dev, cap));
  st.cap = cap;
  return MB200_OK;
}

// Brings `src` (host or device, `bytes`) to the device; returns the device pointer to use.
int stage_in(mb200_scene *s, mb200_scene::Staging &st, const void *src, size_t bytes, const void **dptr) {
  if (is_device_ptr(src)) {
    *dptr = src;
    return MB200_OK;
  }
  int rc = ensure(st, bytes);
  if (rc != MB200_OK) return rc;
  memcpy(st.pinned, src, bytes);
  CU(cudaMemcpyAsync(st.dev, st.pinned, bytes, cudaMemcpyHostToDevice, s->stream));
  *dptr = st.dev;
  return MB200_OK;
}

// Chooses where a kernel should write an output of `bytes`: the caller's device buffer or staging.
int stage_out_begin(mb200_scene::Staging &st, void *dst, size_t bytes, void **dptr, bool *staged) {
  if (is_device_ptr(dst)) {
    *dptr = dst;
    *staged = false;
    return MB200_OK;
  }
  int rc = ensure(st, bytes);
  if (rc != MB200_OK) return rc;
  *dptr = st.dev;
  *staged = true;
  return MB200_OK;
}

int stage_out_enqueue(mb200_scene *s, mb200_scene::Staging &st, size_t bytes) {
  CU(cudaMemcpyAsync(st.pinned, st.dev, bytes, cudaMemcpyDeviceToHost, s->stream));
  return MB200_OK;
}

// the traversal kernels index rays with 32-bit work items (trace_sm.cuh)
constexpr size_t kMaxRaysPerCall = 0xFFFFFFE0ull;

int read_counters(mb200_scene *s, unsigned long long out[4]) {
  CU(cudaMemcpyAsync(out, s->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return MB200_OK;
}

} // namespace

namespace mb200 {
int capi_set_error(int code, const std::string &msg) { return set_err(code, msg); }
} // namespace mb200

extern "C" {

const char *mb200_last_error(void) { return g_err.c_str(); }
const char *mb200_version(void) { return "mallie_b200 0.1 (sm_100a)"; }

int mb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mb200_launches_issued(void) { return mb200::launches_issued(); }

// ---------------------------------------------------------------------------- host BVH
void mb200_build_options_default(mb200_build_options *opt) {
  if (!opt) return;
  opt->cost_taabb = 0.2;
  opt->min_leaf_primitives = 16;
  opt->max_tree_depth = 256;
  opt->bin_size = 64;
}

int mb200_bvh_build(mb200_bvh **out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                    const mb200_build_options *opt) {
  if (!out) return set_err(MB200_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if ((nfaces && (!vertices || !faces))) return set_err(MB200_ERR_INVALID_ARG, "null mesh array");
  mb200_build_options o;
  mb200_build_options_default(&o);
  if (opt) o = *opt;
  mb200_bvh *b = new mb200_bvh;
  std::string err;
  if (!mb200::build_bvh(b->bvh, vertices, nverts, faces, nfaces, o, &err)) {
    delete b;
    return set_err(MB200_ERR_INVALID_ARG, err);
  }
  *out = b;
  return MB200_OK;
}

int mb200_bvh_build_device(mb200_bvh **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                           size_t nfaces, const mb200_build_options *opt) {
  if (!out) return set_err(MB200_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if ((nfaces && (!vertices || !faces))) return set_err(MB200_ERR_INVALID_ARG, "null mesh array");
  mb200_build_options o;
  mb200_build_options_default(&o);
  if (opt) o = *opt;
  mb200_bvh *b = new mb200_bvh;
  std::string err;
  bool cuda_failure = false;
  if (!mb200::build_bvh_device(b->bvh, device, vertices, nverts, faces, nfaces, o, &err, &cuda_failure)) {
    delete b;
    return set_err(cuda_failure ? MB200_ERR_CUDA : MB200_ERR_INVALID_ARG, err);
  }
  *out = b;
  return MB200_OK;
}

int mb200_bvh_load(mb200_bvh **out, const char *path) {
  if (!out || !path) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  mb200_bvh *b = new mb200_bvh;
  std::string err;
  if (!mb200::load_bvh(b->bvh, path, &err)) {
    delete b;
    return set_err(MB200_ERR_IO, err);
  }
  *out = b;
  return MB200_OK;
}

int mb200_bvh_dump(const mb200_bvh *bvh, const char *path) {
  if (!bvh || !path) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  std::string err;
  if (!mb200::dump_bvh(bvh->bvh, path, &err)) return set_err(MB200_ERR_IO, err);
  return MB200_OK;
}

size_t mb200_bvh_num_nodes(const mb200_bvh *bvh) { return bvh ? bvh->bvh.nodes.size() : 0; }
size_t mb200_bvh_num_indices(const mb200_bvh *bvh) { return bvh ? bvh->bvh.indices.size() : 0; }
const mb200_bvh_node *mb200_bvh_nodes(const mb200_bvh *bvh) { return bvh ? bvh->bvh.nodes.data() : nullptr; }
const uint32_t *mb200_bvh_indices(const mb200_bvh *bvh) { return bvh ? bvh->bvh.indices.data() : nullptr; }
int mb200_bvh_stats(const mb200_bvh *bvh, mb200_build_stats *out) {
  if (!bvh || !out) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  *out = bvh->bvh.stats;
  return MB200_OK;
}
void mb200_bvh_destroy(mb200_bvh *bvh) { delete bvh; }

int mb200_bvh_device_layout(const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                            const uint32_t *material_ids, const mb200_bvh_node *nodes, size_t nnodes,
                            const uint32_t *indices, size_t nindices, mb200_layout_info *info, void *pair_nodes_out,
                            void *tri_records_out) {
  if (!info) return set_err(MB200_ERR_INVALID_ARG, "info is null");
  memset(info, 0, sizeof(*info));
  mb200::Relayout r;
  std::string err;
  const int rc = mb200::relayout_bvh(r, vertices, nverts, faces, nfaces, material_ids, nodes, nnodes, indices, nindices, &err);
  if (rc != MB200_OK) return set_err(rc, err);
  info->num_pair_nodes = r.pairs.size();
  info->num_tri_records = r.f32 ? r.tris32.size() : r.tris64.size();
  info->tri_record_bytes = r.f32 ? (uint32_t)sizeof(mb200::TriRecordF32) : (uint32_t)sizeof(mb200::TriRecordF64);
  info->root_ref = r.root_ref, info->root_cnt = r.root_cnt;
  info->depth = r.depth, info->empty = r.empty ? 1 : 0;
  if (pair_nodes_out && !r.pairs.empty()) memcpy(pair_nodes_out, r.pairs.data(), r.pairs.size() * sizeof(mb200::PairNode));
  if (tri_records_out && info->num_tri_records)
    memcpy(tri_records_out, r.f32 ? (const void *)r.tris32.data() : (const void *)r.tris64.data(),
           (size_t)info->num_tri_records * info->tri_record_bytes);
  return MB200_OK;
}

// ---------------------------------------------------------------------------- mesh + config
static int load_mesh(mb200_mesh **out, const char *path, bool eson) {
  if (!out || !path) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  mb200_mesh *m = new (std::nothrow) mb200_mesh();
  if (!m) return set_err(MB200_ERR_OUT_OF_MEMORY, "out of host memory");
  std::string err;
  const bool ok = eson ? mb200::load_eson(m->mesh, path, &err) : mb200::load_obj(m->mesh, path, &err);
  if (!ok) {
    delete m;
    return set_err(MB200_ERR_IO, err);
  }
  *out = m;
  return MB200_OK;
}
int mb200_mesh_load_obj(mb200_mesh **out, const char *path) { return load_mesh(out, path, false); }
int mb200_mesh_load_eson(mb200_mesh **out, const char *path) { return load_mesh(out, path, true); }
int mb200_mesh_transform(mb200_mesh *mesh, double scene_scale, int scene_fit) {
  if (!mesh) return set_err(MB200_ERR_INVALID_ARG, "null mesh");
  mb200::apply_scene_transform(mesh->mesh.vertices.data(), mesh->mesh.vertices.size() / 3, scene_scale,
                               scene_fit != 0, false);
  return MB200_OK;
}
size_t mb200_mesh_num_vertices(const mb200_mesh *m) { return m ? m->mesh.vertices.size() / 3 : 0; }
size_t mb200_mesh_num_faces(const mb200_mesh *m) { return m ? m->mesh.faces.size() / 3 : 0; }
const double *mb200_mesh_vertices(const mb200_mesh *m) { return m ? m->mesh.vertices.data() : nullptr; }
const uint32_t *mb200_mesh_faces(const mb200_mesh *m) { return m ? m->mesh.faces.data() : nullptr; }
const uint32_t *mb200_mesh_material_ids(const mb200_mesh *m) { return m ? m->mesh.material_ids.data() : nullptr; }
const double *mb200_mesh_fv_normals(const mb200_mesh *m) {
  return (m && !m->mesh.normals.empty()) ? m->mesh.normals.data() : nullptr;
}
const double *mb200_mesh_fv_uvs(const mb200_mesh *m) {
  return (m && !m->mesh.uvs.empty()) ? m->mesh.uvs.data() : nullptr;
}
void mb200_mesh_destroy(mb200_mesh *m) { delete m; }

static void config_to_pod(const mallie::RenderConfig &c, mb200_config *o) {
  o->fov = c.fov, o->width = c.width, o->height = c.height;
  for (int k = 0; k < 3; k++)
    o->eye[k] = c.eye[k], o->lookat[k] = c.lookat[k], o->up[k] = c.up[k], o->light[k] = c.light[k];
  for (int k = 0; k < 4; k++) o->quat[k] = c.quat[k];
  o->scene_scale = c.scene_scale, o->scene_fit = c.scene_fit, o->plane = c.plane;
  o->num_passes = c.num_passes, o->num_photons = c.num_photons;
  snprintf(o->obj_filename, sizeof(o->obj_filename), "%s", c.obj_filename.c_str());
  snprintf(o->eson_filename, sizeof(o->eson_filename), "%s", c.eson_filename.c_str());
  snprintf(o->magicavoxel_filename, sizeof(o->magicavoxel_filename), "%s", c.magicavoxel_filename.c_str());
  snprintf(o->material_filename, sizeof(o->material_filename), "%s", c.material_filename.c_str());
  o->max_path_length = c.max_path_length, o->shader = c.shader, o->device = c.device, o->num_gpus = c.num_gpus;
}
static void pod_to_config(const mb200_config *o, mallie::RenderConfig &c) {
  c.fov = o->fov, c.width = o->width, c.height = o->height;
  for (int k = 0; k < 3; k++)
    c.eye[k] = o->eye[k], c.lookat[k] = o->lookat[k], c.up[k] = o->up[k], c.light[k] = o->light[k];
  for (int k = 0; k < 4; k++) c.quat[k] = o->quat[k];
  c.scene_scale = o->scene_scale, c.scene_fit = o->scene_fit != 0, c.plane = o->plane != 0;
  c.num_passes = o->num_passes, c.num_photons = o->num_photons;
  c.obj_filename = o->obj_filename, c.eson_filename = o->eson_filename;
  c.magicavoxel_filename = o->magicavoxel_filename, c.material_filename = o->material_filename;
  c.max_path_length = o->max_path_length, c.shader = o->shader, c.device = o->device, c.num_gpus = o->num_gpus;
}
void mb200_config_default(mb200_config *cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  config_to_pod(mallie::RenderConfig(), cfg);
}
int mb200_config_load(mb200_config *cfg, const char *path, const char *json_text) {
  if (!cfg || (path == nullptr) == (json_text == nullptr))
    return set_err(MB200_ERR_INVALID_ARG, "need a config and exactly one of path / json_text");
  cfg->obj_filename[sizeof(cfg->obj_filename) - 1] = cfg->eson_filename[sizeof(cfg->eson_filename) - 1] = 0;
  cfg->magicavoxel_filename[sizeof(cfg->magicavoxel_filename) - 1] = 0;
  cfg->material_filename[sizeof(cfg->material_filename) - 1] = 0;
  mallie::RenderConfig c;
  pod_to_config(cfg, c);
  const bool ok = path ? mallie::LoadJSONConfig(c, path) : mallie::LoadJSONConfigFromString(c, json_text);
  if (!ok)
    return set_err(MB200_ERR_IO, path ? std::string("cannot read or parse ") + path : std::string("malformed JSON"));
  config_to_pod(c, cfg);
  return MB200_OK;
}

// ---------------------------------------------------------------------------- scene
int mb200_scene_create(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                       size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                       const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices) {
  if (!out) return set_err(MB200_ERR_INVALID_ARG, "out is null");
  std::string err;
  int rc = mb200::scene_create(out, device, vertices, nverts, faces, nfaces, material_ids, fv_normals, fv_uvs, nodes,
                               nnodes, indices, nindices, &err);
  if (rc != MB200_OK) return set_err(rc, err);
  return MB200_OK;
}

int mb200_scene_build(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                      size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                      const mb200_build_options *opt, mb200_bvh **bvh_out) {
  if (!out) return set_err(MB200_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if (bvh_out) *bvh_out = nullptr;
  if ((nfaces && (!vertices || !faces))) return set_err(MB200_ERR_INVALID_ARG, "null mesh array");
  mb200_build_options o;
  mb200_build_options_default(&o);
  if (opt) o = *opt;
  mb200_bvh *b = bvh_out ? new mb200_bvh : nullptr;
  std::string err;
  const int rc = mb200::scene_build_device(out, device, vertices, nverts, faces, nfaces, material_ids, fv_normals, fv_uvs,
                                           o, b ? &b->bvh : nullptr, &err);
  if (rc != MB200_OK) {
    delete b;
    return set_err(rc, err);
  }
  if (bvh_out) *bvh_out = b;
  return MB200_OK;
}

int mb200_scene_layout(mb200_scene *scene, mb200_layout_info *info, void *pair_nodes_out, void *tri_records_out) {
  if (!scene || !info) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  const mb200::SceneView &v = scene->view;
  memset(info, 0, sizeof(*info));
  info->num_pair_nodes = v.num_pair_nodes;
  info->num_tri_records = v.num_tris;
  info->tri_record_bytes = v.tri_f32 ? (uint32_t)sizeof(mb200::TriRecordF32) : (uint32_t)sizeof(mb200::TriRecordF64);
  info->root_ref = v.root_ref, info->root_cnt = v.root_cnt;
  info->depth = scene->tree_depth, info->empty = v.empty;
  if (v.empty || (!pair_nodes_out && !tri_records_out)) return MB200_OK;
  cudaError_t e = cudaSetDevice(scene->device);
  if (e == cudaSuccess) e = cudaStreamSynchronize(scene->stream);
  if (e == cudaSuccess && pair_nodes_out && v.num_pair_nodes)
    e = cudaMemcpy(pair_nodes_out, v.nodes, (size_t)v.num_pair_nodes * sizeof(mb200::PairNode), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && tri_records_out && v.num_tris)
    e = cudaMemcpy(tri_records_out, v.tris, (size_t)v.num_tris * info->tri_record_bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_err(MB200_ERR_CUDA, std::string("scene layout download: ") + cudaGetErrorString(e));
  return MB200_OK;
}

int mb200_scene_clone(mb200_scene **out, mb200_scene *src, int device) {
  if (!out || !src) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  std::string err;
  const int rc = mb200::scene_clone(out, src, device, &err);
  if (rc != MB200_OK) return set_err(rc, err);
  return MB200_OK;
}

void mb200_scene_destroy(mb200_scene *scene) { mb200::scene_destroy(scene); }

int mb200_scene_bounds(const mb200_scene *scene, double bmin[3], double bmax[3]) {
  if (!scene || !bmin || !bmax) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (scene->view.empty) return set_err(MB200_ERR_INVALID_ARG, "empty scene has no bounds");
  for (int k = 0; k < 3; k++) bmin[k] = scene->root_bmin[k], bmax[k] = scene->root_bmax[k];
  return MB200_OK;
}

size_t mb200_scene_device_bytes(const mb200_scene *scene) { return scene ? scene->device_bytes : 0; }
void *mb200_scene_stream(const mb200_scene *scene) { return scene ? (void *)scene->stream : nullptr; }
int mb200_scene_device(const mb200_scene *scene) { return scene ? scene->device : -1; }
int mb200_scene_uses_f32_vertices(const mb200_scene *scene) { return scene ? scene->view.tri_f32 : 0; }

int mb200_scene_timing(mb200_scene *scene, int enable) {
  if (!scene) return set_err(MB200_ERR_INVALID_ARG, "null scene");
  CU(cudaSetDevice(scene->device));
  CU(cudaStreamSynchronize(scene->stream));
  double ms[mb200::kKClasses] = {0};
  unsigned long long n[mb200::kKClasses] = {0};
  scene->timer.collect(ms, n); // drop what was pending
  scene->timer.enabled = enable != 0;
  return MB200_OK;
}

int mb200_scene_kernel_times(mb200_scene *scene, mb200_kernel_times *out) {
  if (!scene || !out) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  CU(cudaSetDevice(scene->device));
  CU(cudaStreamSynchronize(scene->stream));
  double ms[mb200::kKClasses] = {0};
  unsigned long long n[mb200::kKClasses] = {0};
  if (scene->pipe.aux) CU(cudaStreamSynchronize(scene->pipe.aux));
  scene->timer.collect(ms, n, &out->trace_union_ms);
  out->camera_trace_ms = ms[mb200::kKCameraTrace], out->camera_trace_launches = n[mb200::kKCameraTrace];
  out->shadow_trace_ms = ms[mb200::kKShadowTrace], out->shadow_trace_launches = n[mb200::kKShadowTrace];
  out->bounce_trace_ms = ms[mb200::kKBounceTrace], out->bounce_trace_launches = n[mb200::kKBounceTrace];
  out->shade_ms = ms[mb200::kKShade], out->shade_launches = n[mb200::kKShade];
  out->resolve_ms = ms[mb200::kKResolve], out->resolve_launches = n[mb200::kKResolve];
  out->query_trace_ms = ms[mb200::kKQueryTrace], out->query_trace_launches = n[mb200::kKQueryTrace];
  return MB200_OK;
}

int mb200_scene_synchronize(mb200_scene *scene) {
  if (!scene) return set_err(MB200_ERR_INVALID_ARG, "scene is null");
  CU(cudaSetDevice(scene->device));
  CU(cudaStreamSynchronize(scene->stream));
  return MB200_OK;
}

// ---------------------------------------------------------------------------- queries
int mb200_trace_closest_async(mb200_scene *s, const mb200_ray *d_rays, size_t n, mb200_hit *d_hits) {
  if (!s || (n && (!d_rays || !d_hits))) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (n > kMaxRaysPerCall) return set_err(MB200_ERR_INVALID_ARG, "more than 0xFFFFFFE0 rays in one call");
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  CU(mb200::launch_trace_closest(s->view, s->stack_cap, d_rays, n, d_hits, s->d_work, nullptr, s->stream, &s->timer));
  return MB200_OK;
}

int mb200_trace_closest(mb200_scene *s, const mb200_ray *rays, size_t n, mb200_hit *hits, mb200_counters *counters) {
  if (!s || (n && (!rays || !hits))) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (n > kMaxRaysPerCall) return set_err(MB200_ERR_INVALID_ARG, "more than 0xFFFFFFE0 rays in one call");
  if (counters) memset(counters, 0, sizeof(*counters));
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  const void *d_rays;
  void *d_hits;
  bool staged;
  int rc;
  if ((rc = stage_in(s, s->in0, rays, n * sizeof(mb200_ray), &d_rays)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->out0, hits, n * sizeof(mb200_hit), &d_hits, &staged)) != MB200_OK) return rc;
  if (counters) CU(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), s->stream));
  CU(mb200::launch_trace_closest(s->view, s->stack_cap, (const mb200_ray *)d_rays, n, (mb200_hit *)d_hits, s->d_work,
                                 counters ? s->d_counters : nullptr, s->stream, &s->timer));
  if (staged && (rc = stage_out_enqueue(s, s->out0, n * sizeof(mb200_hit))) != MB200_OK) return rc;
  if (counters) {
    unsigned long long c[4];
    if ((rc = read_counters(s, c)) != MB200_OK) return rc;
    counters->nodes_tested = c[0], counters->tris_tested = c[1], counters->rays = c[2], counters->max_stack = c[3];
  }
  CU(cudaStreamSynchronize(s->stream));
  if (staged) memcpy(hits, s->out0.pinned, n * sizeof(mb200_hit));
  return MB200_OK;
}

int mb200_trace_closest_full(mb200_scene *s, const mb200_ray *rays, size_t n, mb200_isect *isects,
                             uint8_t *hit_mask) {
  if (!s || (n && (!rays || !isects))) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (n > kMaxRaysPerCall) return set_err(MB200_ERR_INVALID_ARG, "more than 0xFFFFFFE0 rays in one call");
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  const void *d_rays;
  void *d_is, *d_mask = nullptr;
  bool st_is, st_mask = false;
  int rc;
  if ((rc = stage_in(s, s->in0, rays, n * sizeof(mb200_ray), &d_rays)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->out0, isects, n * sizeof(mb200_isect), &d_is, &st_is)) != MB200_OK) return rc;
  if (hit_mask && (rc = stage_out_begin(s->out1, hit_mask, n, &d_mask, &st_mask)) != MB200_OK) return rc;
  // K2 into scratch hit records, then K3 (BuildIntersection) expands them to the 184-byte records
  CU(mb200::frame_scratch_reserve(s->hit_scratch, n * sizeof(mb200_hit), s->stream));
  mb200_hit *d_hits = (mb200_hit *)s->hit_scratch.base;
  CU(mb200::launch_trace_closest(s->view, s->stack_cap, (const mb200_ray *)d_rays, n, d_hits, s->d_work, nullptr,
                                 s->stream, &s->timer));
  CU(mb200::launch_build_isects(s->view, (const mb200_ray *)d_rays, d_hits, n, (mb200_isect *)d_is,
                                (unsigned char *)d_mask, s->stream));
  if (st_is && (rc = stage_out_enqueue(s, s->out0, n * sizeof(mb200_isect))) != MB200_OK) return rc;
  if (st_mask && (rc = stage_out_enqueue(s, s->out1, n)) != MB200_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  if (st_is) memcpy(isects, s->out0.pinned, n * sizeof(mb200_isect));
  if (st_mask) memcpy(hit_mask, s->out1.pinned, n);
  return MB200_OK;
}

int mb200_trace_occluded(mb200_scene *s, const mb200_ray *rays, const double *tmax, size_t n, uint8_t *occluded,
                         mb200_counters *counters) {
  if (!s || (n && (!rays || !tmax || !occluded))) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (n > kMaxRaysPerCall) return set_err(MB200_ERR_INVALID_ARG, "more than 0xFFFFFFE0 rays in one call");
  if (counters) memset(counters, 0, sizeof(*counters));
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  const void *d_rays, *d_tmax;
  void *d_occ;
  bool staged;
  int rc;
  if ((rc = stage_in(s, s->in0, rays, n * sizeof(mb200_ray), &d_rays)) != MB200_OK) return rc;
  if ((rc = stage_in(s, s->in1, tmax, n * sizeof(double), &d_tmax)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->out0, occluded, n, &d_occ, &staged)) != MB200_OK) return rc;
  if (counters) CU(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), s->stream));
  CU(mb200::launch_trace_occluded(s->view, s->stack_cap, (const mb200_ray *)d_rays, (const double *)d_tmax, n,
                                  (unsigned char *)d_occ, s->d_work, counters ? s->d_counters : nullptr, s->stream,
                                  &s->timer));
  if (staged && (rc = stage_out_enqueue(s, s->out0, n)) != MB200_OK) return rc;
  if (counters) {
    unsigned long long c[4];
    if ((rc = read_counters(s, c)) != MB200_OK) return rc;
    counters->nodes_tested = c[0], counters->tris_tested = c[1], counters->rays = c[2], counters->max_stack = c[3];
  }
  CU(cudaStreamSynchronize(s->stream));
  if (staged) memcpy(occluded, s->out0.pinned, n);
  return MB200_OK;
}

// ---------------------------------------------------------------------------- camera
int mb200_generate_rays(mb200_scene *s, const mb200_camera_frame *frame, const double *px, const double *py, size_t n,
                        mb200_ray *rays) {
  if (!s || !frame || (n && (!px || !py || !rays))) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  const void *d_px, *d_py;
  void *d_rays;
  bool staged;
  int rc;
  if ((rc = stage_in(s, s->in0, px, n * sizeof(double), &d_px)) != MB200_OK) return rc;
  if ((rc = stage_in(s, s->in1, py, n * sizeof(double), &d_py)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->out0, rays, n * sizeof(mb200_ray), &d_rays, &staged)) != MB200_OK) return rc;
  CU(mb200::launch_generate_rays(*frame, (const double *)d_px, (const double *)d_py, n, (mb200_ray *)d_rays,
                                 s->stream));
  if (staged && (rc = stage_out_enqueue(s, s->out0, n * sizeof(mb200_ray))) != MB200_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  if (staged) memcpy(rays, s->out0.pinned, n * sizeof(mb200_ray));
  return MB200_OK;
}

int mb200_generate_rays_env(mb200_scene *s, const double origin[3], int width, int height, const double *px,
                            const double *py, size_t n, int stereo, mb200_ray *rays) {
  if (!s || !origin || width <= 0 || height <= 0 || (n && (!px || !py || !rays)))
    return set_err(MB200_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  const void *d_px, *d_py;
  void *d_rays;
  bool staged;
  int rc;
  if ((rc = stage_in(s, s->in0, px, n * sizeof(double), &d_px)) != MB200_OK) return rc;
  if ((rc = stage_in(s, s->in1, py, n * sizeof(double), &d_py)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->out0, rays, n * sizeof(mb200_ray), &d_rays, &staged)) != MB200_OK) return rc;
  CU(mb200::launch_generate_rays_env(origin, width, height, stereo, (const double *)d_px, (const double *)d_py, n,
                                     (mb200_ray *)d_rays, s->stream));
  if (staged && (rc = stage_out_enqueue(s, s->out0, n * sizeof(mb200_ray))) != MB200_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  if (staged) memcpy(rays, s->out0.pinned, n * sizeof(mb200_ray));
  return MB200_OK;
}

int mb200_generate_rays_grid(mb200_scene *s, const mb200_camera_frame *frame, int x0, int y0, int x1, int y1,
                             mb200_ray *rays) {
  if (!s || !frame || !rays || x1 < x0 || y1 < y0) return set_err(MB200_ERR_INVALID_ARG, "bad argument");
  const size_t n = (size_t)(x1 - x0) * (size_t)(y1 - y0);
  if (n == 0) return MB200_OK;
  CU(cudaSetDevice(s->device));
  void *d_rays;
  bool staged;
  int rc;
  if ((rc = stage_out_begin(s->out0, rays, n * sizeof(mb200_ray), &d_rays, &staged)) != MB200_OK) return rc;
  CU(mb200::launch_generate_grid(*frame, x0, y0, x1 - x0, y1 - y0, (mb200_ray *)d_rays, s->stream));
  if (staged && (rc = stage_out_enqueue(s, s->out0, n * sizeof(mb200_ray))) != MB200_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  if (staged) memcpy(rays, s->out0.pinned, n * sizeof(mb200_ray));
  return MB200_OK;
}

// ---------------------------------------------------------------------------- frame
void mb200_render_params_default(mb200_render_params *p, int width, int height) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->width = width, p->height = height;
  p->x0 = 0, p->y0 = 0, p->x1 = width, p->y1 = height;
  p->max_path_length = 16;
  p->jitter = 1;
  p->shader = MB200_SHADER_PATHTRACE;
  p->light[0] = 0.0, p->light[1] = 20.0, p->light[2] = 0.0;
}

// gPlaneObject.set(0, 1, 0, -(zmin - zsize * 0.0001f)) with float zmin/zsize (render.cc:620-627)
void mb200_plane_from_bounds(const double bmin[3], const double bmax[3], float abcd[4]) {
  const float zmin = (float)bmin[1];
  const float zsize = (float)(bmax[1] - bmin[1]);
  abcd[0] = 0, abcd[1] = 1, abcd[2] = 0;
  abcd[3] = -(zmin - zsize * 0.0001f);
}

// mode 0: one pass, image overwritten, count++ (Render);  1: accumulate (image +=, count +=);
// mode 2: fresh frame (image = sum, count = N; nothing is read from the caller's buffers).
static int render_common(mb200_scene *s, const mb200_render_params *p, int num_passes, int mode, float *image,
                         int *count, mb200_render_stats *stats) {
  if (!s || !p || !image || !count) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (p->width <= 0 || p->height <= 0 || p->x0 < 0 || p->y0 < 0 || p->x1 > p->width || p->y1 > p->height ||
      p->x0 > p->x1 || p->y0 > p->y1 || p->max_path_length < 1 || num_passes < 1)
    return set_err(MB200_ERR_INVALID_ARG, "bad render parameters");
  if (p->shader < 0 || p->shader > MB200_SHADER_PATHTRACE_ENV) return set_err(MB200_ERR_INVALID_ARG, "unknown shader");
  if (p->camera_mode < 0 || p->camera_mode > MB200_CAMERA_ENV_STEREO)
    return set_err(MB200_ERR_INVALID_ARG, "unknown camera mode");
  if (p->band_rows < 0 || (p->band_rows > 0 && (p->band_rows % 4 != 0 || p->band_count < 1 || p->band_index < 0 ||
                                                 p->band_index >= p->band_count)))
    return set_err(MB200_ERR_INVALID_ARG, "bad band parameters (band_rows must be a multiple of 4)");
  if (p->pixel_step < 0 || (p->pixel_step > 1 && p->band_rows > 0))
    return set_err(MB200_ERR_INVALID_ARG, "pixel_step must be >= 0 and cannot be combined with row bands");
  if (stats) memset(stats, 0, sizeof(*stats));
  CU(cudaSetDevice(s->device));
  const bool compact = p->band_rows > 0 && p->band_compact;
  const size_t rows = compact ? (size_t)mb200_band_local_rows(p) : (size_t)p->height;
  const size_t npix = (size_t)p->width * rows;
  if (npix == 0) return MB200_OK;
  const size_t img_bytes = npix * 3 * sizeof(float), cnt_bytes = npix * sizeof(int);
  const HostKind img_kind = classify(image), cnt_kind = classify(count);
  float *d_img = image;
  int *d_cnt = count;
  int rc;
  // whole-buffer coverage: every pixel of the buffer is produced by this call
  const bool covers_all = compact || (p->band_rows == 0 && p->x0 == 0 && p->y0 == 0 && p->x1 == p->width &&
                                      p->y1 == p->height);
  const bool need_img_in = (mode == 1) || !covers_all;  // pixels outside the tile must survive
  const bool need_cnt_in = (mode != 2) || !covers_all;
  if (img_kind != kDevice) {
    if ((rc = ensure(s->out0, img_bytes, img_kind == kPageable)) != MB200_OK) return rc;
    d_img = (float *)s->out0.dev;
    if (need_img_in) {
      const void *src = image;
      if (img_kind == kPageable) memcpy(s->out0.pinned, image, img_bytes), src = s->out0.pinned;
      CU(cudaMemcpyAsync(d_img, src, img_bytes, cudaMemcpyHostToDevice, s->stream));
    }
  }
  if (cnt_kind != kDevice) {
    if ((rc = ensure(s->out1, cnt_bytes, cnt_kind == kPageable)) != MB200_OK) return rc;
    d_cnt = (int *)s->out1.dev;
    if (need_cnt_in) {
      const void *src = count;
      if (cnt_kind == kPageable) memcpy(s->out1.pinned, count, cnt_bytes), src = s->out1.pinned;
      CU(cudaMemcpyAsync(d_cnt, src, cnt_bytes, cudaMemcpyHostToDevice, s->stream));
    }
  }
  if (stats) CU(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), s->stream));
  // A host framebuffer is copied back in row chunks: the frame is cut into batches by rows, so the first rows are final
  // while the later ones are still traced, and their device -> host copy runs on the pipe's copy stream meanwhile.
  mb200::FrameChunks chunks;
  const bool to_host = img_kind != kDevice || cnt_kind != kDevice;
  CU(mb200::launch_frame(s->view, s->stack_cap, *p, num_passes, mode, d_img, d_cnt, s->frame_scratch,
                         stats ? s->d_counters : nullptr, s->stream, &s->timer, &s->pipe, to_host ? &chunks : nullptr));
  // A fresh frame that covers the buffer has count == num_passes everywhere (every step-th pixel aside): the host
  // writes it itself while the GPU renders instead of waiting for 4 more bytes per pixel over PCIe.
  const bool count_is_constant = mode == 2 && covers_all && p->pixel_step <= 1 && cnt_kind != kDevice;
  const bool copy_img = img_kind != kDevice, copy_cnt = cnt_kind != kDevice && !count_is_constant;
  void *img_dst = img_kind == kPinned ? (void *)image : s->out0.pinned;
  void *cnt_dst = cnt_kind == kPinned ? (void *)count : s->out1.pinned;
  bool chunked = false;
  if (to_host && chunks.n > 0 && s->pipe.copy) {
    chunked = true;
    const size_t W = (size_t)p->width;
    for (int c = 0; c < chunks.n; c++) {
      const size_t r0 = (size_t)chunks.row0[c], nr = (size_t)(chunks.row1[c] - chunks.row0[c]);
      CU(cudaStreamWaitEvent(s->pipe.copy, chunks.done_a[c], 0));
      if (chunks.done_b[c]) CU(cudaStreamWaitEvent(s->pipe.copy, chunks.done_b[c], 0));
      if (copy_img)
        CU(cudaMemcpyAsync((char *)img_dst + r0 * W * 3 * sizeof(float), (const char *)d_img + r0 * W * 3 * sizeof(float),
                           nr * W * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->pipe.copy));
      if (copy_cnt)
        CU(cudaMemcpyAsync((char *)cnt_dst + r0 * W * sizeof(int), (const char *)d_cnt + r0 * W * sizeof(int),
                           nr * W * sizeof(int), cudaMemcpyDeviceToHost, s->pipe.copy));
    }
  } else {
    if (copy_img) CU(cudaMemcpyAsync(img_dst, d_img, img_bytes, cudaMemcpyDeviceToHost, s->stream));
    if (copy_cnt) CU(cudaMemcpyAsync(cnt_dst, d_cnt, cnt_bytes, cudaMemcpyDeviceToHost, s->stream));
  }
  if (count_is_constant) std::fill_n(count, npix, num_passes);
  unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (stats || to_host) {
    if (stats) CU(cudaMemcpyAsync(c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (chunked) CU(cudaStreamSynchronize(s->pipe.copy));
  }
  if (img_kind == kPageable) memcpy(image, s->out0.pinned, img_bytes);
  if (cnt_kind == kPageable && !count_is_constant) memcpy(count, s->out1.pinned, cnt_bytes);
  if (stats) {
    stats->primary_rays = c[0], stats->bounce_rays = c[1], stats->shadow_rays = c[2], stats->zombie_segments = c[3];
    stats->camera_nodes_tested = c[4], stats->camera_tris_tested = c[5];
    stats->shadow_nodes_tested = c[6], stats->shadow_tris_tested = c[7];
  }
  return MB200_OK;
}

// Can `dev` address memory of `root`?  Enables peer access on first use.
static bool peer_ok(int dev, int root) {
  static std::atomic<int> state[64][64]; // 0 unknown, 1 yes, 2 no; hosts may drive frames from several threads
  if (dev == root) return true;
  if (dev < 0 || root < 0 || dev >= 64 || root >= 64) return false;
  int st = state[dev][root].load(std::memory_order_acquire);
  if (st == 0) { // two threads asking at once both probe; enabling twice is harmless (AlreadyEnabled)
    int can = 0;
    st = 2;
    if (cudaDeviceCanAccessPeer(&can, dev, root) == cudaSuccess && can) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(dev);
      const cudaError_t e = cudaDeviceEnablePeerAccess(root, 0);
      if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) st = 1;
      cudaGetLastError();
      cudaSetDevice(cur);
    }
    state[dev][root].store(st, std::memory_order_release);
  }
  return st == 1;
}

int mb200_render_frame_multi(mb200_scene *const *scenes, int n, const mb200_render_params *p, int num_passes,
                             int band_rows, float *image, int *count, mb200_render_stats *stats) {
  if (!scenes || n < 1 || !p || !image || !count) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  for (int g = 0; g < n; g++)
    if (!scenes[g]) return set_err(MB200_ERR_INVALID_ARG, "null scene");
  if (n == 1) return render_common(scenes[0], p, num_passes, 2, image, count, stats);
  if (p->width <= 0 || p->height <= 0 || p->x0 != 0 || p->y0 != 0 || p->x1 != p->width || p->y1 != p->height ||
      p->band_rows != 0 || p->pixel_step > 1 || p->max_path_length < 1 || num_passes < 1 || band_rows < 4 ||
      band_rows % 4 != 0 || p->shader < 0 || p->shader > MB200_SHADER_PATHTRACE_ENV || p->camera_mode < 0 ||
      p->camera_mode > MB200_CAMERA_ENV_STEREO)
    return set_err(MB200_ERR_INVALID_ARG,
                   "multi-GPU frames need whole-image parameters without bands / step and band_rows = 4k");
  if (n > 64) return set_err(MB200_ERR_INVALID_ARG, "too many scenes");
  for (int g = 0; g < n; g++)
    for (int h = 0; h < g; h++)
      if (scenes[g]->device == scenes[h]->device)
        return set_err(MB200_ERR_INVALID_ARG, "every scene must live on its own GPU");
  if (stats) memset(stats, 0, sizeof(*stats));
  mb200_scene *root = scenes[0];
  const size_t W = (size_t)p->width, H = (size_t)p->height;
  const size_t img_bytes = W * H * 3 * sizeof(float), cnt_bytes = W * H * sizeof(int);
  const HostKind img_kind = classify(image), cnt_kind = classify(count);
  CU(cudaSetDevice(root->device));
  float *d_img = image;
  int *d_cnt = count;
  int rc;
  if (img_kind != kDevice) {
    if ((rc = ensure(root->out0, img_bytes, img_kind == kPageable)) != MB200_OK) return rc;
    d_img = (float *)root->out0.dev;
  }
  if (cnt_kind != kDevice) {
    if ((rc = ensure(root->out1, cnt_bytes, cnt_kind == kPageable)) != MB200_OK) return rc;
    d_cnt = (int *)root->out1.dev;
  }
  cudaEvent_t done[64];
  int ndone = 0;
  auto cleanup = [&]() {
    for (int k = 0; k < ndone; k++) cudaEventDestroy(done[k]);
  };
  for (int g = 0; g < n; g++) {
    mb200_scene *s = scenes[g];
    cudaError_t e = cudaSetDevice(s->device);
    mb200_render_params pg = *p;
    pg.band_rows = band_rows, pg.band_count = n, pg.band_index = g, pg.band_compact = 0;
    float *t_img = d_img;
    int *t_cnt = d_cnt;
    const bool direct = peer_ok(s->device, root->device);
    size_t rows = 0;
    if (e == cudaSuccess && !direct) { // local compact band buffer, copied to the root afterwards
      pg.band_compact = 1;
      rows = (size_t)mb200_band_local_rows(&pg);
      if (ensure(s->in0, rows * W * 3 * sizeof(float), false) != MB200_OK || ensure(s->in1, rows * W * sizeof(int), false) != MB200_OK) {
        cleanup();
        return set_err(MB200_ERR_OUT_OF_MEMORY, "mb200_render_frame_multi: band buffers: " + g_err);
      }
      t_img = (float *)s->in0.dev, t_cnt = (int *)s->in1.dev;
    }
    if (e == cudaSuccess && stats) e = cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), s->stream);
    if (e == cudaSuccess)
      e = mb200::launch_frame(s->view, s->stack_cap, pg, num_passes, 2, t_img, t_cnt, s->frame_scratch,
                              stats ? s->d_counters : nullptr, s->stream, &s->timer, &s->pipe);
    if (e == cudaSuccess && !direct) {
      size_t local = 0;
      const size_t nbands = (H + band_rows - 1) / band_rows;
      for (size_t b = (size_t)g; b < nbands && e == cudaSuccess; b += (size_t)n) {
        const size_t y = b * band_rows, r = (y + band_rows <= H) ? (size_t)band_rows : H - y;
        e = cudaMemcpyPeerAsync(d_img + y * W * 3, root->device, t_img + local * W * 3, s->device, r * W * 3 * sizeof(float), s->stream);
        if (e == cudaSuccess)
          e = cudaMemcpyPeerAsync(d_cnt + y * W, root->device, t_cnt + local * W, s->device, r * W * sizeof(int), s->stream);
        local += r;
      }
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[ndone], cudaEventDisableTiming);
    if (e == cudaSuccess) {
      ndone++;
      e = cudaEventRecord(done[ndone - 1], s->stream);
    }
    if (e != cudaSuccess) {
      cleanup();
      return cuda_err(e, "mb200_render_frame_multi: enqueue");
    }
  }
  cudaError_t e = cudaSetDevice(root->device);
  for (int k = 0; k < ndone && e == cudaSuccess; k++) e = cudaStreamWaitEvent(root->stream, done[k], 0);
  if (e == cudaSuccess && img_kind != kDevice)
    e = cudaMemcpyAsync(img_kind == kPinned ? (void *)image : root->out0.pinned, d_img, img_bytes, cudaMemcpyDeviceToHost, root->stream);
  if (e == cudaSuccess && cnt_kind != kDevice)
    e = cudaMemcpyAsync(cnt_kind == kPinned ? (void *)count : root->out1.pinned, d_cnt, cnt_bytes, cudaMemcpyDeviceToHost, root->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(root->stream);
  cleanup();
  if (e != cudaSuccess) return cuda_err(e, "mb200_render_frame_multi: gather");
  if (img_kind == kPageable) memcpy(image, root->out0.pinned, img_bytes);
  if (cnt_kind == kPageable) memcpy(count, root->out1.pinned, cnt_bytes);
  if (stats) {
    for (int g = 0; g < n; g++) {
      unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      CU(cudaSetDevice(scenes[g]->device));
      CU(cudaMemcpy(c, scenes[g]->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
      stats->primary_rays += c[0], stats->bounce_rays += c[1], stats->shadow_rays += c[2], stats->zombie_segments += c[3];
      stats->camera_nodes_tested += c[4], stats->camera_tris_tested += c[5];
      stats->shadow_nodes_tested += c[6], stats->shadow_tris_tested += c[7];
    }
    CU(cudaSetDevice(root->device));
  }
  return MB200_OK;
}

int mb200_resolve_ldr(mb200_scene *s, const float *image, const int *count, int width, int height, int mode,
                      unsigned char *out) {
  if (!s || !image || !count || !out || width <= 0 || height <= 0) return set_err(MB200_ERR_INVALID_ARG, "bad argument");
  if (mode != MB200_LDR_RGB8_LINEAR && mode != MB200_LDR_BGRA8_GAMMA22) return set_err(MB200_ERR_INVALID_ARG, "unknown LDR mode");
  CU(cudaSetDevice(s->device));
  const size_t npix = (size_t)width * height;
  const size_t out_bytes = npix * (mode == MB200_LDR_RGB8_LINEAR ? 3 : 4);
  const void *d_img, *d_cnt;
  void *d_out;
  bool staged;
  int rc;
  // the staging slots a render call with host buffers used for its outputs are free again: out0 / out1 hold device
  // copies of a host frame, in0 the 8-bit result
  if ((rc = stage_in(s, s->out0, image, npix * 3 * sizeof(float), &d_img)) != MB200_OK) return rc;
  if ((rc = stage_in(s, s->out1, count, npix * sizeof(int), &d_cnt)) != MB200_OK) return rc;
  if ((rc = stage_out_begin(s->in0, out, out_bytes, &d_out, &staged)) != MB200_OK) return rc;
  CU(mb200::launch_resolve_ldr((const float *)d_img, (const int *)d_cnt, npix, mode, (unsigned char *)d_out, s->stream));
  if (!staged) return MB200_OK; // device destination: enqueue-only
  if ((rc = stage_out_enqueue(s, s->in0, out_bytes)) != MB200_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  memcpy(out, s->in0.pinned, out_bytes);
  return MB200_OK;
}

int mb200_render_frame_ldr(mb200_scene *s, const mb200_render_params *p, int num_passes, int mode, unsigned char *out,
                           mb200_render_stats *stats) {
  if (!s || !p || !out) return set_err(MB200_ERR_INVALID_ARG, "null argument");
  if (mode != MB200_LDR_RGB8_LINEAR && mode != MB200_LDR_BGRA8_GAMMA22) return set_err(MB200_ERR_INVALID_ARG, "unknown LDR mode");
  if (p->width <= 0 || p->height <= 0 || p->x0 != 0 || p->y0 != 0 || p->x1 != p->width || p->y1 != p->height || p->band_rows != 0)
    return set_err(MB200_ERR_INVALID_ARG, "mb200_render_frame_ldr renders whole images");
  CU(cudaSetDevice(s->device));
  const size_t npix = (size_t)p->width * p->height;
  int rc;
  // the float frame and the counts live in the scene's device staging and never leave the GPU
  if ((rc = ensure(s->out0, npix * 3 * sizeof(float), false)) != MB200_OK) return rc;
  if ((rc = ensure(s->out1, npix * sizeof(int), false)) != MB200_OK) return rc;
  if ((rc = render_common(s, p, num_passes, 2, (float *)s->out0.dev, (int *)s->out1.dev, stats)) != MB200_OK) return rc;
  const size_t out_bytes = npix * (mode == MB200_LDR_RGB8_LINEAR ? 3 : 4);
  void *d_out;
  bool staged;
  if ((rc = stage_out_begin(s->in0, out, out_bytes, &d_out, &staged)) != MB200_OK) return rc;
  CU(mb200::launch_resolve_ldr((const float *)s->out0.dev, (const int *)s->out1.dev, npix, mode, (unsigned char *)d_out, s->stream));
  if (!staged) return MB200_OK;
  const bool pinned = classify(out) == kPinned;
  CU(cudaMemcpyAsync(pinned ? (void *)out : s->in0.pinned, d_out, out_bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (!pinned) memcpy(out, s->in0.pinned, out_bytes);
  return MB200_OK;
}

int mb200_band_local_rows(const mb200_render_params *p) {
  if (!p) return 0;
  if (p->band_rows <= 0) return p->y1 - p->y0;
  return mb200::band_rows_owned(p->y1 - p->y0, p->band_rows, p->band_count, p->band_index);
}

int mb200_render_pass(mb200_scene *scene, const mb200_render_params *params, float *image, int *count,
                      mb200_render_stats *stats) {
  return render_common(scene, params, 1, 0, image, count, stats);
}

int mb200_render_accumulate(mb200_scene *scene, const mb200_render_params *params, int num_passes, float *accum,
                            int *count, mb200_render_stats *stats) {
  return render_common(scene, params, num_passes, 1, accum, count, stats);
}

int mb200_render_frame(mb200_scene *scene, const mb200_render_params *params, int num_passes, float *image,
                       int *count, mb200_render_stats *stats) {
  return render_common(scene, params, num_passes, 2, image, count, stats);
}

} // extern "C"

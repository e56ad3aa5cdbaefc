/* mallie_b200.h -- C ABI of the B200-native Mallie render hot path.
 *
 * Drop-in boundary.  Mallie has no FFI layer of its own; the
 * seam it uses to swap its ray-tracing backend is the ENABLE_EMBREE #ifdef in
 * Scene (scene.h:70-76, scene.cc:172-221 build, :254-311 trace, :318-320 bbox).
 * This header is what a `#ifdef ENABLE_B200` branch at the same three places
 * binds to (see INTEGRATION.md for the exact patch), plus the frame-level entry
 * that replaces the OpenMP scanline loop of mallie::Render (render.cc:657-698).
 *
 * Conventions
 *  - plain C, pointers + sizes, no C++/torch types; every function returns
 *    MB200_OK (0) or a negative mb200_status and never aborts; the message for
 *    the calling thread's last failure is mb200_last_error().
 *  - "rays", "hits", "image" ... buffers may be HOST or DEVICE pointers (the
 *    library asks cudaPointerGetAttributes); host buffers are staged through
 *    pinned memory inside the call.  Calls block until host-visible results are
 *    readable.  When every buffer of a call is a DEVICE pointer and no host-side
 *    output (counters / stats) is requested, mb200_render_* and the *_async
 *    variants only enqueue on the scene's stream: order later work on
 *    mb200_scene_stream() or call mb200_scene_synchronize().
 *  - inputs are borrowed for the duration of the call only.
 *  - one scene lives on one GPU; use one scene per GPU for multi-GPU (tiles).
 *  - All arithmetic that decides a result is IEEE double in the reference's
 *    operation order without FMA contraction: hit records are bit-identical
 *    to BVHAccel::Traverse (bvh_accel.cc:773-844).
 */
#ifndef MALLIE_B200_H_
#define MALLIE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MB200_OK = 0,
  MB200_ERR_INVALID_ARG = -1,  /* null pointer, bad size, malformed BVH        */
  MB200_ERR_CUDA = -2,         /* a CUDA runtime call failed                    */
  MB200_ERR_NO_DEVICE = -3,    /* no usable sm_100 GPU / bad device ordinal     */
  MB200_ERR_OUT_OF_MEMORY = -4,
  MB200_ERR_IO = -5,           /* file could not be read / written              */
  MB200_ERR_UNSUPPORTED = -6
} mb200_status;

/* -------------------------------------------------------------------------
 * POD records.  Layouts are the leading fields of the reference structs so a
 * Mallie caller can pass its own arrays without conversion.
 * ---------------------------------------------------------------------- */

/* struct Ray (common.h:78-83): org, dir.  invDir/dirSign are recomputed by
 * Traverse itself (bvh_accel.cc:787-797) and are not part of the ABI.  48 B. */
typedef struct { double org[3]; double dir[3]; } mb200_ray;

/* Head of struct Intersection (intersection.h:6-11).  32 B.
 * Miss: t = DBL_MAX, u = v = 0, faceID = 0xFFFFFFFF, materialID = 0xFFFFFFFF. */
typedef struct { double t, u, v; uint32_t faceID, materialID; } mb200_hit;

/* Full struct Intersection (intersection.h:6-24), 184 B, as filled by
 * BuildIntersection (bvh_accel.cc:699-769).  tangent/binormal are never written
 * by the BVH path (left zero here).  On a miss only t,u,v,faceID are defined. */
typedef struct {
  double t, u, v;
  uint32_t faceID, materialID;
  uint32_t f0, f1, f2, pad_;
  double position[3];
  double geometricNormal[3];
  double normal[3];
  double tangent[3];
  double binormal[3];
  double texcoord[2];
} mb200_isect;

/* class BVHNode (bvh_accel.h:10-29), 64 B: what BVHAccel::GetNodes() returns
 * and BVHAccel::Dump writes. */
typedef struct {
  double bmin[3];
  double bmax[3];
  int32_t flag;     /* 1 = leaf, 0 = branch */
  int32_t axis;     /* branch: split axis; leaf: ignored */
  uint32_t data[2]; /* branch: child0, child1; leaf: ntris, first index */
} mb200_bvh_node;

/* struct BVHBuildOptions (bvh_accel.h:32-42). */
typedef struct {
  double cost_taabb;       /* 0.2 */
  int min_leaf_primitives; /* 16  */
  int max_tree_depth;      /* 256 */
  int bin_size;            /* 64  */
} mb200_build_options;

/* struct BVHBuildStatistics (bvh_accel.h:45-52). */
typedef struct { int max_tree_depth, num_leaf_nodes, num_branch_nodes; } mb200_build_stats;

/* Per-call traversal counters (optional outputs). */
typedef struct {
  uint64_t nodes_tested; /* box tests == nodes popped by the reference loop (bvh_accel.cc:805-811) */
  uint64_t tris_tested;  /* triangles run through TriangleIsect (bvh_accel.cc:656-694)            */
  uint64_t rays;         /* rays traced                                                            */
  uint64_t max_stack;    /* deepest traversal stack seen                                           */
} mb200_counters;

typedef struct mb200_bvh mb200_bvh;     /* host-side BVH (reference layout)           */
typedef struct mb200_scene mb200_scene; /* device-resident scene: BVH + mesh, one GPU */

/* -------------------------------------------------------------------------
 * Library
 * ---------------------------------------------------------------------- */
const char *mb200_last_error(void);
const char *mb200_version(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int mb200_device_count(void);
/* Kernel launches this library has issued in this process (monitoring; bench.py's gpu_launches). */
int mb200_launches_issued(void);

/* -------------------------------------------------------------------------
 * Host BVH: replaces BVHAccel::Build / Dump / Load (bvh_accel.cc:445-544).
 * The tree is bit-identical to the reference builder's (same node order,
 * same index permutation) so faceID tie-breaks agree.
 * ---------------------------------------------------------------------- */
void mb200_build_options_default(mb200_build_options *opt);
int mb200_bvh_build(mb200_bvh **out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                    const mb200_build_options *opt /* NULL = defaults */);
/* BVHAccel::Build on GPU `device`: the same tree as mb200_bvh_build, bit for bit (nodes, bounds, index order),
 * grown level by level with one pass per level over all open nodes (mallie_b200/csrc/device/bvh_build_gpu.cu).
 * Needs min_leaf_primitives >= 2.  MB200_ERR_CUDA when no usable device. */
int mb200_bvh_build_device(mb200_bvh **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                           size_t nfaces, const mb200_build_options *opt /* NULL = defaults */);
int mb200_bvh_load(mb200_bvh **out, const char *path);       /* BVHAccel::Load  */
int mb200_bvh_dump(const mb200_bvh *bvh, const char *path);  /* BVHAccel::Dump  */
size_t mb200_bvh_num_nodes(const mb200_bvh *bvh);
size_t mb200_bvh_num_indices(const mb200_bvh *bvh);
const mb200_bvh_node *mb200_bvh_nodes(const mb200_bvh *bvh); /* BVHAccel::GetNodes   */
const uint32_t *mb200_bvh_indices(const mb200_bvh *bvh);     /* BVHAccel::GetIndices */
int mb200_bvh_stats(const mb200_bvh *bvh, mb200_build_stats *out); /* BVHAccel::GetStatistics */
void mb200_bvh_destroy(mb200_bvh *bvh);

/* The device layout mb200_scene_create would upload for this mesh + reference-layout BVH (host only, no GPU
 * needed): the tree is validated exactly as by mb200_scene_create (MB200_ERR_INVALID_ARG + message for a
 * malformed one), re-laid out as 128-byte pair nodes (both children's boxes in the parent) and per-leaf-order
 * triangle records (mallie_b200/csrc/device/layout.h).  pair_nodes_out / tri_records_out may be NULL (sizes
 * only); otherwise they receive num_pair_nodes * 128 and num_tri_records * tri_record_bytes bytes. */
typedef struct {
  uint64_t num_pair_nodes, num_tri_records;
  uint32_t tri_record_bytes; /* 48: float-exact vertices, 80: double p0 + edges */
  uint32_t root_ref, root_cnt; /* as a pair node's ref / cnt; cnt == 0xFFFFFFFF: the root is a branch */
  int32_t depth, empty;
} mb200_layout_info;
int mb200_bvh_device_layout(const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                            const uint32_t *material_ids, const mb200_bvh_node *nodes, size_t nnodes,
                            const uint32_t *indices, size_t nindices, mb200_layout_info *info,
                            void *pair_nodes_out, void *tri_records_out);

/* -------------------------------------------------------------------------
 * Mesh ingestion and configuration: what Scene::Init does before the BVH build
 * (scene.cc:66-170) and LoadJSONConfig (main.cc:98-205).  Host only.
 * ---------------------------------------------------------------------- */
typedef struct mb200_mesh mb200_mesh; /* host mesh: the arrays struct Mesh (mesh.h:7-18) points at */
/* MeshLoader::LoadObj (importers/mesh_loader.cc:26-210 over tiny_obj_loader.cc): same vertex / face
 * numbering, float-parsed positions, fan triangulation, material ids, face-varying normals and uvs. */
int mb200_mesh_load_obj(mb200_mesh **out, const char *path);
/* MeshLoader::LoadESON (importers/mesh_loader.cc:212-310). */
int mb200_mesh_load_eson(mb200_mesh **out, const char *path);
/* Scene::Init's vertex transform (scene.cc:112-170): scene_fit != 0 maps the bounding box to [-1,1]^3,
 * otherwise vertices *= scene_scale. */
int mb200_mesh_transform(mb200_mesh *mesh, double scene_scale, int scene_fit);
size_t mb200_mesh_num_vertices(const mb200_mesh *mesh);
size_t mb200_mesh_num_faces(const mb200_mesh *mesh);
const double *mb200_mesh_vertices(const mb200_mesh *mesh);       /* [3*nv] */
const uint32_t *mb200_mesh_faces(const mb200_mesh *mesh);        /* [3*nf] */
const uint32_t *mb200_mesh_material_ids(const mb200_mesh *mesh); /* [nf]   */
const double *mb200_mesh_fv_normals(const mb200_mesh *mesh);     /* [9*nf] or NULL */
const double *mb200_mesh_fv_uvs(const mb200_mesh *mesh);         /* [6*nf] or NULL */
void mb200_mesh_destroy(mb200_mesh *mesh);

/* struct RenderConfig (render.h:11-49) as a POD; defaults are RenderConfig()'s (render.h:33-48). */
typedef struct {
  double fov;
  int width, height;
  double eye[3], lookat[3], up[3], quat[4];
  double scene_scale;
  int scene_fit, plane;
  int num_passes, num_photons;
  char obj_filename[1024], eson_filename[1024], magicavoxel_filename[1024], material_filename[1024];
  /* additions; absent keys keep the reference's behaviour */
  int max_path_length; /* "max_path_length", default 16 (kMaxPathLength, render.cc:52) */
  int shader;          /* "shader": "pathtrace" (default) | "primary_shadow" | "primary"   */
  double light[3];     /* "light"                                                            */
  int device, num_gpus; /* "device", "gpus"                                                  */
} mb200_config;
void mb200_config_default(mb200_config *cfg);
/* LoadJSONConfig (main.cc:98-205).  Exactly one of path / json_text is non-NULL.  Keys that are absent
 * leave *cfg untouched (call mb200_config_default first). */
int mb200_config_load(mb200_config *cfg, const char *path, const char *json_text);

/* -------------------------------------------------------------------------
 * Device scene: replaces the accel_ member of Scene (scene.h:75) and what
 * Scene::Init does after loading the mesh (scene.cc:224-230).
 * vertices [3*nverts] f64, faces [3*nfaces] u32 as in struct Mesh (mesh.h:7-18);
 * material_ids [nfaces], fv_normals [9*nfaces], fv_uvs [6*nfaces] may be NULL.
 * nodes/indices: the reference-layout BVH (from mb200_bvh_* or from Mallie's own
 * BVHAccel::GetNodes()/GetIndices()).  The library validates the tree, re-lays it
 * out for the GPU and uploads it; nothing is retained from the caller's arrays.
 * ---------------------------------------------------------------------- */
int mb200_scene_create(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                       size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                       const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices);
/* BVHAccel::Build (bvh_accel.cc:445) and the scene upload in one step, entirely on GPU `device`: the tree is grown
 * there (as mb200_bvh_build_device) and the traversal layout is written from it on the device; only the mesh is
 * uploaded.  The scene equals mb200_scene_create(mb200_bvh_build(...)) byte for byte.  bvh_out (may be NULL)
 * receives the reference-layout tree (for BVHAccel::GetNodes / Dump or replicas on other GPUs). */
int mb200_scene_build(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                      size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                      const mb200_build_options *opt /* NULL = defaults */, mb200_bvh **bvh_out);
/* The traversal layout resident on the device (same meaning and sizes as mb200_bvh_device_layout): for checking
 * a device-built scene against a host-built one.  The output pointers may be NULL (sizes only). */
int mb200_scene_layout(mb200_scene *scene, mb200_layout_info *info, void *pair_nodes_out, void *tri_records_out);
/* A replica of `src` on GPU `device` (the per-GPU copies of the multi-GPU frame): the resident arrays are copied
 * device to device, over NVLink when the GPUs are peers; nothing is rebuilt or re-laid out on the host. */
int mb200_scene_clone(mb200_scene **out, mb200_scene *src, int device);
void mb200_scene_destroy(mb200_scene *scene);
/* Scene::BoundingBox (scene.cc:317-333): root node bounds. */
int mb200_scene_bounds(const mb200_scene *scene, double bmin[3], double bmax[3]);
/* Bytes resident in HBM for this scene, and the CUDA stream (cudaStream_t) work is enqueued on. */
size_t mb200_scene_device_bytes(const mb200_scene *scene);
void *mb200_scene_stream(const mb200_scene *scene);
int mb200_scene_device(const mb200_scene *scene);
/* 1 when every vertex coordinate is exactly float-representable and the compact
 * fp32 triangle records are in use (still widened to double before any arithmetic). */
int mb200_scene_uses_f32_vertices(const mb200_scene *scene);

/* Per-kernel device timing (monitoring; what the reference's timerutil around Render() is to the CPU path,
 * render.cc:630-707, at kernel granularity).  While enabled, every kernel the scene launches is bracketed by
 * CUDA events on the scene's stream.  mb200_scene_kernel_times synchronises the stream, returns the
 * milliseconds and launch counts accumulated since the last call (per kernel class) and resets them. */
typedef struct {
  /* per class: length of the union of the launches' time spans (consecutive batches of a frame run on two
   * streams, so launches of a class can overlap each other and launches of other classes) */
  double camera_trace_ms, shadow_trace_ms, bounce_trace_ms, shade_ms, resolve_ms, query_trace_ms;
  uint64_t camera_trace_launches, shadow_trace_launches, bounce_trace_launches, shade_launches, resolve_launches,
      query_trace_launches;
  double trace_union_ms; /* time during which at least one camera / shadow / bounce traversal launch was in flight */
} mb200_kernel_times;
int mb200_scene_timing(mb200_scene *scene, int enable);
int mb200_scene_kernel_times(mb200_scene *scene, mb200_kernel_times *out);

/* Measured ceilings of GPU `device` (diagnostics for roofline reports; a few hundred milliseconds, 2 GiB of
 * scratch): FP64 pipe lane-operations / s (independent DADD / DMUL chains; the kernels are built without FMA
 * contraction), bytes / s of 256-bit loads over an L2-resident 64 MB buffer, bytes / s (read + write) of a 1 GiB
 * copy.  What the reference's timerutil is to its render loop (render.cc:630-707), for the device's limits. */
typedef struct {
  double fp64_lane_ops_per_s, l2_read_bytes_per_s, hbm_copy_bytes_per_s;
  int sm_count;
} mb200_peaks;
int mb200_probe_peaks(int device, mb200_peaks *out);

/* -------------------------------------------------------------------------
 * Queries: replace bool Scene::Trace(Intersection&, Ray&) (scene.cc:253-315) ->
 * BVHAccel::Traverse (bvh_accel.cc:773-844), batched.
 * ---------------------------------------------------------------------- */
/* Closest hit; hits[i] is the 32-byte record head.  counters may be NULL. */
int mb200_trace_closest(mb200_scene *scene, const mb200_ray *rays, size_t n, mb200_hit *hits,
                        mb200_counters *counters);
/* Closest hit + BuildIntersection: full 184-byte records. hit_mask (n bytes) may be NULL. */
int mb200_trace_closest_full(mb200_scene *scene, const mb200_ray *rays, size_t n, mb200_isect *isects,
                             uint8_t *hit_mask);
/* Occlusion (any hit): occluded[i] = 1 iff closest-hit Traverse would return t < tmax[i]
 * (the shadow-ray query the reference leaves as an empty block, render.cc:425-426). */
int mb200_trace_occluded(mb200_scene *scene, const mb200_ray *rays, const double *tmax, size_t n, uint8_t *occluded,
                         mb200_counters *counters);
/* Enqueue-only variants for DEVICE buffers (no host synchronisation); pair with
 * mb200_scene_synchronize or CUDA events on mb200_scene_stream(). */
int mb200_trace_closest_async(mb200_scene *scene, const mb200_ray *d_rays, size_t n, mb200_hit *d_hits);
int mb200_scene_synchronize(mb200_scene *scene);

/* -------------------------------------------------------------------------
 * Camera: Camera::BuildCameraFrame (camera.cc:40-220) on the host,
 * Camera::GenerateRay (camera.cc:222-240) on the device.
 * ---------------------------------------------------------------------- */
typedef struct { double origin[3], corner[3], du[3], dv[3]; } mb200_camera_frame;
int mb200_camera_frame_build(mb200_camera_frame *out, const double eye[3], const double lookat[3], const double up[3],
                             double fov, const double quat[4], int width, int height);
/* rays[i] = GenerateRay(px[i], py[i]); px/py/rays host or device. */
int mb200_generate_rays(mb200_scene *scene, const mb200_camera_frame *frame, const double *px, const double *py,
                        size_t n, mb200_ray *rays);
/* rays[i] = GenerateEnvRay(px[i], py[i]) or, stereo != 0, GenerateStereoEnvRay(px[i], py[i]) for a width x height
 * panorama seen from origin (camera.cc:242-329).  sin / cos / fmod / atan2 are CUDA's: directions agree with the
 * reference's glibc results to a few ulp, not bit for bit. */
int mb200_generate_rays_env(mb200_scene *scene, const double origin[3], int width, int height, const double *px,
                            const double *py, size_t n, int stereo, mb200_ray *rays);
/* Un-jittered primary rays GenerateRay((double)x,(double)y) for the tile [x0,x1)x[y0,y1), row-major. */
int mb200_generate_rays_grid(mb200_scene *scene, const mb200_camera_frame *frame, int x0, int y0, int x1, int y1,
                             mb200_ray *rays);

/* -------------------------------------------------------------------------
 * Frame: replaces the body of mallie::Render (render.cc:593-708) for one pass.
 * ---------------------------------------------------------------------- */
typedef enum {
  MB200_SHADER_PATHTRACE = 0,      /* PathTrace, render.cc:381-456 */
  MB200_SHADER_PRIMARY_SHADOW = 1, /* primary closest hit + one shadow ray to `light` */
  MB200_SHADER_PRIMARY_ONLY = 2,   /* primary closest hit; radiance = |normal| visualisation-free: hit ? 1 : 0 */
  MB200_SHADER_PATHTRACE_ENV = 3   /* PathTraceEnv, render.cc:518-590: PathTrace without the plane and without the
                                      material attenuation (miss term 0.5 / pathLength); used by RenderPanoramic */
} mb200_shader;

typedef enum {
  MB200_CAMERA_PINHOLE = 0,   /* Camera::GenerateRay,          camera.cc:222-240 */
  MB200_CAMERA_ENV = 1,       /* Camera::GenerateEnvRay,       camera.cc:242-257 (equirectangular panorama) */
  MB200_CAMERA_ENV_STEREO = 2 /* Camera::GenerateStereoEnvRay, camera.cc:259-329 (top / bottom stereo pair) */
} mb200_camera_mode;

typedef struct {
  int width, height;          /* full image size (RenderConfig::width/height)                        */
  int x0, y0, x1, y1;         /* tile rendered by this call (whole image: 0,0,width,height)           */
  mb200_camera_frame frame;
  int use_plane;              /* RenderConfig::plane                                                  */
  float plane[4];             /* Plane::set(a,b,c,d), render.cc:620-627 (mb200_plane_from_bounds)     */
  int max_path_length;        /* kMaxPathLength, render.cc:52 (16)                                    */
  uint32_t pass;              /* sample index: seeds the per-pixel RNG stream                         */
  int jitter;                 /* 1 = PathTrace's [-0.5,0.5) jitter (render.cc:388-391); 0 = pixel coords as is */
  int shader;                 /* mb200_shader                                                         */
  double light[3];            /* MB200_SHADER_PRIMARY_SHADOW                                          */
  /* Multi-GPU row-band interleave.  band_rows == 0: disabled.  Otherwise the tile's rows are
   * cut into bands of band_rows scanlines (a multiple of 4) and this call renders only bands
   * b with b % band_count == band_index; with band_compact != 0 the image/count buffers hold just those
   * rows, packed band after band (float[3*width*mb200_band_local_rows()], the NCCL gather send buffer). */
  int band_rows, band_count, band_index, band_compact;
  /* Render()'s `step` argument (render.cc:657-698).  0 or 1: every pixel.  step > 1 (coarse preview): one
   * sample at every step-th pixel of every step-th row, copied over its step x step block (clipped to the
   * tile), and count += 3 for every pixel of the block -- the reference increments it inside the k < 3
   * colour loop (render.cc:689-693).  Not combinable with band_rows. */
  int pixel_step;
  int camera_mode;            /* mb200_camera_mode; the panorama cameras use frame.origin and width / height only */
} mb200_render_params;

typedef struct {
  uint64_t primary_rays;  /* closest-hit queries for camera rays               */
  uint64_t bounce_rays;   /* closest-hit queries for path continuation         */
  uint64_t shadow_rays;   /* occlusion queries                                 */
  uint64_t zombie_segments; /* post-escape segments resolved in closed form, NOT counted as rays */
  /* Work the traversal kernel did for the frame, counted as mb200_counters counts it (box tests / triangles run
   * through TriangleIsect), split by ray kind.  Filled for MB200_SHADER_PRIMARY_SHADOW / _PRIMARY_ONLY frames;
   * 0 for the path-tracing shaders.  The shadow figures are what the any-hit walk really visited
   * (it stops at the first t < tmax), i.e. at most the closest-hit Traverse counts that define the query. */
  uint64_t camera_nodes_tested, camera_tris_tested, shadow_nodes_tested, shadow_tris_tested;
} mb200_render_stats;

void mb200_render_params_default(mb200_render_params *p, int width, int height);
void mb200_plane_from_bounds(const double bmin[3], const double bmax[3], float abcd[4]);
/* image: float[3*width*height] RGB row-major, FULL-image indexing (only the tile's pixels are written,
 * overwritten not accumulated, as render.cc:673-675); count: int[width*height], ++ per tile pixel
 * (render.cc:677-679).  Host or device pointers.  stats may be NULL. */
int mb200_render_pass(mb200_scene *scene, const mb200_render_params *params, float *image, int *count,
                      mb200_render_stats *stats);
/* N passes accumulated on the device: accum[p] += pass image, count[p] += N; what the SDL render thread does
 * with AccumImage (main_sdl.cc:572-606).  accum: float[3*W*H], count: int[W*H], host or device. */
int mb200_render_accumulate(mb200_scene *scene, const mb200_render_params *params, int num_passes, float *accum,
                            int *count, mb200_render_stats *stats);
/* A whole frame of num_passes samples per pixel: image = sum of the passes, count = num_passes, both
 * OVERWRITTEN for the tile's pixels (what DoMainConsole does with zeroed buffers and one pass,
 * main_console.cc:57-75, generalised to N passes); nothing is read from image/count, so host
 * buffers cost one device->host copy only. */
int mb200_render_frame(mb200_scene *scene, const mb200_render_params *params, int num_passes, float *image,
                       int *count, mb200_render_stats *stats);
/* One frame over several GPUs from ONE host thread (the single-process form of SURVEY.md §8e; one process per
 * GPU uses mb200_render_frame with band_* and an NCCL gather instead, bench.py).  scenes[g] is a replica of the
 * same scene on its own GPU; the image rows are cut into bands of band_rows scanlines (a multiple of 4), band b
 * is rendered by scenes[b % num_scenes], and the frame is assembled in scenes[0]'s GPU: with peer access every
 * GPU's resolve kernel stores its rows straight into that framebuffer over NVLink (no copy, no collective);
 * without it each GPU renders into a local band buffer that is copied peer-to-peer afterwards.  params must
 * describe the whole image (x0 = y0 = 0, x1 = width, y1 = height, band_rows = 0, pixel_step <= 1).
 * image / count: as mb200_render_frame (device pointers must belong to scenes[0]'s GPU).  stats: sums. */
int mb200_render_frame_multi(mb200_scene *const *scenes, int num_scenes, const mb200_render_params *params,
                             int num_passes, int band_rows, float *image, int *count, mb200_render_stats *stats);
/* Output resolve on the device: the accumulated frame (float[3*W*H]) and its per-pixel sample counts (int[W*H]) ->
 * 8-bit pixels, as the reference's front ends produce them on the host.  With the frame still on the device (what
 * mb200_render_frame / _accumulate leave there) a host receives W*H*3 or W*H*4 bytes instead of W*H*16.
 * image / count / out: host or device pointers; a device `out` makes the call enqueue-only. */
typedef enum {
  MB200_LDR_RGB8_LINEAR = 0,   /* HDRToLDR, main_console.cc:25-43: out[3p+c] = clamp((int)(in / count * 255.5))            */
  MB200_LDR_BGRA8_GAMMA22 = 1  /* Display, main_sdl.cc:156-165,420-477: BGRA, powf(in / count, 1 / 2.2f) * 255.5, A = 255 */
} mb200_ldr_mode;
int mb200_resolve_ldr(mb200_scene *scene, const float *image, const int *count, int width, int height, int mode,
                      unsigned char *out);

/* mb200_render_frame + mb200_resolve_ldr without the float frame ever leaving the GPU: what DoMainConsole does with
 * Render + HDRToLDR (main_console.cc:57-75), for num_passes samples per pixel.  params must describe the whole image. */
int mb200_render_frame_ldr(mb200_scene *scene, const mb200_render_params *params, int num_passes, int mode,
                           unsigned char *out, mb200_render_stats *stats);

/* Rows of the image a banded call owns (== y1-y0 when bands are disabled). */
int mb200_band_local_rows(const mb200_render_params *params);

/* -------------------------------------------------------------------------
 * One process per GPU: every rank holds a replica of the scene, renders the row bands
 * b with b % ranks == rank, and ONE NCCL all-gather per frame assembles the framebuffer; the rows are put in
 * place by this library's own kernel on the scene's stream.  The reference has no counterpart (its MPI is an
 * init / finalize stub, main.cc:213-236); what it replaces is DoMainConsole's single-process Render call
 * (main_console.cc:57-75) when the host is launched once per GPU.
 * NCCL is bound at run time (dlopen of libnccl.so.2); without it these return MB200_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------- */
typedef struct mb200_comm mb200_comm;
#define MB200_COMM_ID_BYTES 128 /* == NCCL_UNIQUE_ID_BYTES */
/* ncclGetUniqueId: rank 0 creates the id and hands it to the other ranks by any means it has (file, socket, MPI). */
int mb200_comm_unique_id(unsigned char id[MB200_COMM_ID_BYTES]);
/* ncclCommInitRank on the scene's GPU (collective: every rank calls it). */
int mb200_comm_init(mb200_comm **out, mb200_scene *scene, int nranks, int rank, const unsigned char id[MB200_COMM_ID_BYTES]);
/* Uses a communicator the host already has (nccl_comm is its ncclComm_t); it is not destroyed by mb200_comm_destroy. */
int mb200_comm_adopt(mb200_comm **out, mb200_scene *scene, void *nccl_comm);
int mb200_comm_size(const mb200_comm *comm);
int mb200_comm_rank(const mb200_comm *comm);
/* How the last gathered frame was exchanged: 1 = rows stored straight into every rank's frame buffer over peer memory
 * (k_exchange_rows; one node, every rank can map every other rank's buffer), 0 = ncclAllGather + row placement. */
int mb200_comm_exchange_path(const mb200_comm *comm);
void mb200_comm_destroy(mb200_comm *comm);
/* The exchange of the bands (peer-memory stores when every rank can map every other rank's frame buffer -- decided once
 * per frame size by a collective set-up on the first call -- else ncclAllGather + row placement; MB200_GATHER=nccl forces
 * the latter).  d_bands: this rank's compact band buffer (device, float[channels*width*
 * mb200_band_local_rows()], what a band_compact render call wrote); image: float[channels*width*height], device
 * pointer (enqueue-only on the scene's stream) or host pointer (blocks until readable), or NULL on a rank that
 * does not need the assembled frame (it still takes part in the collective). */
int mb200_gather_framebuffer(mb200_comm *comm, int width, int height, int channels, int band_rows, const float *d_bands,
                             float *image);
/* mb200_render_frame for a frame split over the communicator's ranks: renders this rank's bands of the whole-image
 * `params` straight into the NCCL send buffer, gathers, and delivers the assembled frame to `image` (as above; count
 * receives num_passes everywhere and may be NULL).  stats: this rank's share. */
int mb200_render_frame_gathered(mb200_comm *comm, const mb200_render_params *params, int num_passes, int band_rows,
                                float *image, int *count, mb200_render_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* MALLIE_B200_H_ */

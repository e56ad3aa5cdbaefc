/* oracle/mallie_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of Mallie's hot path (lighttransport/mallie @ 2ec03e06),
 * used ONLY as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product (mallie_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * unmodified reference compiled into oracle/_ref/libmallie_ref.so
 * (tests/test_oracle_vs_reference.py) and against the golden vectors of
 * SURVEY.md App. B (tests/golden/, tests/test_oracle_golden.py).
 */
#ifndef MALLIE_ORACLE_H_
#define MALLIE_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bvh_accel.h:10-29 -- 64 bytes, same field offsets as the reference BVHNode. */
typedef struct {
  double bmin[3];
  double bmax[3];
  int32_t flag; /* 1 = leaf, 0 = branch */
  int32_t axis; /* branch only; the reference leaves it uninitialised in leaves, we store 0 */
  uint32_t data[2]; /* branch: child0, child1; leaf: ntris, first index */
} ora_node;

/* mesh.h:7-18 (only the arrays the path reads) */
typedef struct {
  size_t num_vertices;
  size_t num_faces;
  const double *vertices;      /* [3*nv] */
  const uint32_t *faces;       /* [3*nf] */
  const uint32_t *material_ids; /* [nf] or NULL */
  const double *fv_normals;    /* [9*nf] or NULL */
  const double *fv_uvs;        /* [6*nf] or NULL */
} ora_mesh;

/* head of Intersection, intersection.h:6-11 */
typedef struct {
  double t, u, v;
  uint32_t face_id, material_id;
} ora_hit;

/* full Intersection, intersection.h:6-24 (184 bytes) */
typedef struct {
  double t, u, v;
  uint32_t face_id, material_id;
  uint32_t f0, f1, f2, pad_;
  double position[3];
  double geometric_normal[3];
  double normal[3];
  double tangent[3];
  double binormal[3];
  double texcoord[2];
} ora_isect;

typedef struct ora_bvh ora_bvh;

/* --- builder: bvh_accel.cc:36-482 ------------------------------------- */
ora_bvh *ora_bvh_build(const ora_mesh *mesh, double cost_taabb, int min_leaf, int max_depth, int bin_size);
ora_bvh *ora_bvh_from_arrays(const ora_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices);
void ora_bvh_free(ora_bvh *b);
size_t ora_bvh_num_nodes(const ora_bvh *b);
size_t ora_bvh_num_indices(const ora_bvh *b);
const ora_node *ora_bvh_nodes(const ora_bvh *b);
const uint32_t *ora_bvh_indices(const ora_bvh *b);
void ora_bvh_stats(const ora_bvh *b, int out3[3]); /* maxTreeDepth, numLeafNodes, numBranchNodes */
/* Dump/Load byte format: bvh_accel.cc:484-544 */
int ora_bvh_dump(const ora_bvh *b, const char *path);
ora_bvh *ora_bvh_load(const char *path);

/* --- traversal: bvh_accel.cc:546-844 ---------------------------------- */
/* One ray. Returns 1 on hit. counters (nullable): [0]+=nodes popped, [1]+=tris tested, [2]=max(stack depth). */
int ora_traverse(const ora_bvh *b, const ora_mesh *mesh, const double org[3], const double dir[3],
                 ora_isect *isect, uint64_t counters[3]);
/* Batch, OpenMP over chunks of `row` rays. hits/isects/mask/per-ray counters nullable.
 * per_ray_counts: [2*n] uint16 (nodes, tris) per ray. totals: [3] as in ora_traverse. Returns seconds. */
double ora_trace_batch(const ora_bvh *b, const ora_mesh *mesh, const double *rays, size_t n, ora_hit *hits,
                       ora_isect *isects, uint8_t *mask, uint64_t totals[3], int row, int nthreads);
/* Occlusion oracle: closest-hit Traverse returns t < tmax[i]. */
void ora_occluded_batch(const ora_bvh *b, const ora_mesh *mesh, const double *rays, const double *tmax, size_t n,
                        uint8_t *occluded, int nthreads);

/* --- camera: camera.cc:12-240, matrix.cc:42-216, trackball.cc:272-292 -- */
void ora_camera_frame(const double eye[3], const double lookat[3], const double up[3], double fov,
                      const double quat[4], int width, int height, double origin[3], double corner[3],
                      double du[3], double dv[3]);
void ora_generate_ray(const double origin[3], const double corner[3], const double du[3], const double dv[3],
                      double u, double v, double ray6[6]);
/* Camera::GenerateEnvRay (camera.cc:242-257) / GenerateStereoEnvRay (camera.cc:259-329): equirectangular
 * panorama rays from pixel coordinates (u, v) of a width x height image; stereo != 0 = top/bottom stereo pair. */
void ora_generate_env_ray(const double origin[3], int width, int height, double u, double v, int stereo,
                          double ray6[6]);
void ora_generate_grid(const double origin[3], const double corner[3], const double du[3], const double dv[3],
                       int width, int height, double *rays);

/* --- plane: prim-plane.cc:8-44 ----------------------------------------- */
int ora_plane_intersect(const float abcd[4], const double org[3], const double dir[3], ora_isect *isect);
/* gPlaneObject.set(...) from the scene bbox, render.cc:620-627 */
void ora_plane_from_bbox(const double bmin[3], const double bmax[3], float abcd[4]);

/* --- RNG: render.cc:116-168 -------------------------------------------- */
typedef struct { uint32_t x, y, z, w; } ora_rng;
void ora_rng_seed_reference(ora_rng *r, int tid);          /* init_randomreal */
void ora_rng_seed_pixel(ora_rng *r, uint32_t pixel, uint32_t pass); /* counter-based seeding used on the GPU */
double ora_randomreal(ora_rng *r);

/* --- render: render.cc:381-456, 593-708 -------------------------------- */
typedef struct {
  int width, height;
  double origin[3], corner[3], du[3], dv[3];
  int use_plane;
  float plane[4];
  int max_path_length; /* kMaxPathLength, 16 */
  int rng_mode;        /* 0 = reference sequential stream (tid 0, scanline order; == OMP_NUM_THREADS=1),
                          1 = per-pixel counter seeding ora_rng_seed_pixel(pixel, pass) */
  uint32_t pass;
  int skip_zombies;    /* 1 = do not trace post-escape segments (closed form); same image */
  int shader;          /* 0 = PathTrace (render.cc:381), 1 = primary + shadow (direct light),
                          2 = PathTraceEnv (render.cc:518-590): no plane, no material attenuation */
  double light[3];     /* shader 1 */
  int camera_mode;     /* 0 = Camera::GenerateRay, 1 = GenerateEnvRay, 2 = GenerateStereoEnvRay */
} ora_render_params;
/* image: float[3*W*H] overwritten; count[W*H] incremented; x0..x1,y0..y1 tile (whole image: 0,0,W,H).
 * ray_counts (nullable, accumulated): [0] Trace calls for camera/bounce rays, [1] zombie segments,
 * [2] shadow rays, [3] nodes popped, [4] triangles tested over all those rays (reference traversal order;
 * shadow rays counted as the closest-hit Traverse that defines the occlusion oracle), [5] / [6] the share
 * of [3] / [4] that belongs to the shadow rays. */
void ora_render_pass(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, int x0, int y0, int x1,
                     int y1, float *image, int *count, uint64_t ray_counts[7], int nthreads);
/* Same, also emitting the exact ray set of the pass for the CPU-baseline timing (shader 1 only):
 * primary_rays_out [6*W*H]; shadow_rays_out [7*W*H] = org, dir, tmax (NaN where the camera ray missed). */
void ora_render_pass_ex(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, int x0, int y0, int x1,
                        int y1, float *image, int *count, uint64_t ray_counts[7], int nthreads,
                        double *primary_rays_out, double *shadow_rays_out);

/* RenderPanoramic, render.cc:710-763: image zeroed, then per pixel TEN samples of PathTraceEnv added
 * (float += double), count += 10.  p->camera_mode selects env (1) / stereo (2); p->shader is taken as 2.
 * rng_mode 0 = the reference's sequential stream (OMP_NUM_THREADS=1); rng_mode 1 = per-pixel seeding with
 * pass index p->pass + sample. */
void ora_render_panoramic(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, float *image,
                          int *count, int nthreads);

/* --- output resolve --------------------------------------------------------
 * HDRToLDR + fclamp, main_console.cc:25-43: out[i] = clamp((int)(in[i] / in_count[i / 3] * 255.5)), RGB8.
 * (pinned: the reference function itself is compiled into oracle/_ref through ref_console.cc.) */
void ora_hdr_to_ldr(const float *in, const int *in_count, int width, int height, unsigned char *out);
/* Display + fclamp of the SDL viewer, main_sdl.cc:156-165,420-477: BGRA8, scale = 1.0f / (float)count,
 * clamp((int)(powf(scale * in, 1.0f / 2.2f) * 255.5)), alpha 255.  (restated only: main_sdl.cc needs SDL headers,
 * which this image does not have; powf is glibc's, as in the reference build.) */
void ora_display_bgra(const float *in, const int *counts, int width, int height, unsigned char *out);

/* --- misc ---------------------------------------------------------------- */
uint64_t ora_fnv1a64(const void *data, size_t nbytes, uint64_t seed /* 0 = standard offset basis */);

#ifdef __cplusplus
}
#endif
#endif

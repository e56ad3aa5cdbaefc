/* oracle/mallie_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the Mallie hot path; see mallie_oracle.h for the
 * contract and the parity status (PINNED against oracle/_ref and SURVEY App. B).
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Compile with -ffp-contract=off: the reference is built
 * -O2 -msse2 on x86-64, i.e. plain IEEE double, no fused multiply-add, and all
 * expression orders below are the reference's.
 */
#include "mallie_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* std::numeric_limits<double>::epsilon() * 1024 (bvh_accel.cc:86,158,283,598) */
#define ORA_KEPS (DBL_EPSILON * 1024.0)

static double ora_now(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

/* ======================================================================
 * real3 helpers -- common.h:9-76.  vdot = (a0*b0 + a1*b1) + a2*b2.
 * ==================================================================== */
typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_ptr(const double *p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_scale(v3 a, double f) { return v3_make(a.x * f, a.y * f, a.z * f); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3_cross(v3 a, v3 b) { /* common.h:66-72 */
  return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 v3_normalize(v3 a) { /* common.h:48-57 */
  double len = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
  if (fabs(len) > 1.0e-6) {
    double inv = 1.0 / len;
    a.x *= inv; a.y *= inv; a.z *= inv;
  }
  return a;
}
static inline double v3_get(const v3 *a, int i) { return (&a->x)[i]; }

/* ======================================================================
 * BVH container
 * ==================================================================== */
struct ora_bvh {
  ora_node *nodes;
  size_t nnodes, cap;
  uint32_t *indices;
  size_t nindices;
  int stats[3]; /* maxTreeDepth, numLeafNodes, numBranchNodes */
};

static size_t push_node(ora_bvh *b, const ora_node *n) {
  if (b->nnodes == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 1024;
    b->nodes = (ora_node *)realloc(b->nodes, b->cap * sizeof(ora_node));
  }
  b->nodes[b->nnodes] = *n;
  return b->nnodes++;
}

/* ======================================================================
 * Builder -- bvh_accel.cc:36-482
 * ==================================================================== */
typedef struct {
  const ora_mesh *mesh;
  ora_bvh *bvh;
  double taabb;
  int min_leaf, max_depth, nbins;
  size_t *bins; /* [2][3][nbins]: fresh & zeroed per node, bvh_accel.cc:36-46,378 */
} build_ctx;

/* CalculateSurfaceArea, bvh_accel.cc:50-53 */
static inline double surface_area(const double mn[3], const double mx[3]) {
  double bx = mx[0] - mn[0], by = mx[1] - mn[1], bz = mx[2] - mn[2];
  return 2.0 * (bx * by + by * bz + bz * bx);
}

/* ComputeBoundingBox, bvh_accel.cc:279-315 */
static void compute_bbox(const build_ctx *c, uint32_t l, uint32_t r, double mn[3], double mx[3]) {
  const double *V = c->mesh->vertices;
  const uint32_t *F = c->mesh->faces;
  const uint32_t *I = c->bvh->indices;
  size_t first = I[l];
  for (int k = 0; k < 3; k++) {
    mn[k] = V[3 * (size_t)F[3 * first] + k] - ORA_KEPS;
    mx[k] = V[3 * (size_t)F[3 * first] + k] + ORA_KEPS;
  }
  for (size_t i = l; i < r; i++) {
    size_t tri = I[i];
    for (int j = 0; j < 3; j++) {
      size_t vid = F[3 * tri + j];
      for (int k = 0; k < 3; k++) {
        double lo = V[3 * vid + k] - ORA_KEPS;
        double hi = V[3 * vid + k] + ORA_KEPS;
        if (mn[k] > lo) mn[k] = lo;
        if (mx[k] < hi) mx[k] = hi;
      }
    }
  }
}

/* GetBoundingBoxOfTriangle, bvh_accel.cc:55-80 (std::min/std::max semantics) */
static void tri_bbox(const ora_mesh *m, uint32_t tri, double mn[3], double mx[3]) {
  const double *p0 = &m->vertices[3 * (size_t)m->faces[3 * (size_t)tri + 0]];
  for (int k = 0; k < 3; k++) mn[k] = mx[k] = p0[k];
  for (int j = 1; j < 3; j++) {
    const double *p = &m->vertices[3 * (size_t)m->faces[3 * (size_t)tri + j]];
    for (int k = 0; k < 3; k++) {
      if (p[k] < mn[k]) mn[k] = p[k]; /* std::min(a,b) = (b<a)?b:a */
      if (mx[k] < p[k]) mx[k] = p[k]; /* std::max(a,b) = (a<b)?b:a */
    }
  }
}

/* ContributeBinBuffer, bvh_accel.cc:82-142 */
static void fill_bins(build_ctx *c, const double nmin[3], const double nmax[3], uint32_t l, uint32_t r) {
  const int nb = c->nbins;
  double bin_count = (double)nb;
  double inv[3];
  for (int a = 0; a < 3; a++) {
    double size = nmax[a] - nmin[a];
    inv[a] = (size > ORA_KEPS) ? bin_count / size : 0.0;
  }
  memset(c->bins, 0, sizeof(size_t) * 2 * 3 * (size_t)nb);
  for (size_t i = l; i < r; i++) {
    double tmn[3], tmx[3];
    tri_bbox(c->mesh, c->bvh->indices[i], tmn, tmx);
    for (int a = 0; a < 3; a++) {
      double qmin = (tmn[a] - nmin[a]) * inv[a];
      double qmax = (tmx[a] - nmin[a]) * inv[a];
      size_t imin = (unsigned int)floor(qmin);
      size_t imax = (unsigned int)floor(qmax);
      if ((double)imin >= bin_count) imin = (size_t)(nb - 1);
      if ((double)imax >= bin_count) imax = (size_t)(nb - 1);
      c->bins[0 * (nb * 3) + a * nb + imin] += 1;
      c->bins[1 * (nb * 3) + a * nb + imax] += 1;
    }
  }
}

/* SAH, bvh_accel.cc:144-154 */
static inline double sah_cost(size_t ns1, double left_area, size_t ns2, double right_area, double inv_s,
                              double taabb, double ttri) {
  return 2.0f * taabb + (left_area * inv_s) * (double)(ns1)*ttri + (right_area * inv_s) * (double)(ns2)*ttri;
}

/* FindCutFromBinBuffer, bvh_accel.cc:156-255 */
static void find_cut(const build_ctx *c, const double nmin[3], const double nmax[3], size_t ntris,
                     double cut_pos[3], int *cut_axis) {
  const int nb = c->nbins;
  double ttri = 1.0 - c->taabb;
  double bsize[3], bstep[3], min_cost[3];
  for (int a = 0; a < 3; a++) {
    bsize[a] = nmax[a] - nmin[a];
    bstep[a] = bsize[a] * (1.0 / nb);
  }
  double sa_total = surface_area(nmin, nmax);
  double inv_sa = (sa_total > ORA_KEPS) ? 1.0 / sa_total : 0.0;

  for (int a = 0; a < 3; a++) {
    double best_pos = nmin[a] + 0.5 * bstep[a];
    min_cost[a] = DBL_MAX;
    size_t left = 0, right = ntris;
    double lmin[3], lmax[3], rmin[3], rmax[3];
    for (int k = 0; k < 3; k++) { lmin[k] = rmin[k] = nmin[k]; lmax[k] = rmax[k] = nmax[k]; }
    for (int i = 0; i < nb - 1; ++i) {
      left += c->bins[0 * (3 * nb) + a * nb + i];
      right -= c->bins[1 * (3 * nb) + a * nb + i];
      double pos = nmin[a] + (i + 0.5) * bstep[a];
      lmax[a] = pos;
      rmin[a] = pos;
      double sl = surface_area(lmin, lmax);
      double sr = surface_area(rmin, rmax);
      double cost = sah_cost(left, sl, right, sr, inv_sa, c->taabb, ttri);
      if (cost < min_cost[a]) {
        min_cost[a] = cost;
        best_pos = pos;
      }
    }
    cut_pos[a] = best_pos;
  }
  double cost = min_cost[0];
  *cut_axis = 0;
  if (cost > min_cost[1]) { *cut_axis = 1; cost = min_cost[1]; }
  if (cost > min_cost[2]) { *cut_axis = 2; cost = min_cost[2]; }
}

/* SAHPred, bvh_accel.cc:257-277 */
static inline int sah_pred(const ora_mesh *m, uint32_t tri, int axis, double pos) {
  const uint32_t *f = &m->faces[3 * (size_t)tri];
  double center = m->vertices[3 * (size_t)f[0] + axis] + m->vertices[3 * (size_t)f[1] + axis] +
                  m->vertices[3 * (size_t)f[2] + axis];
  return center < pos * 3.0;
}

/* libstdc++ std::partition for bidirectional iterators (bits/stl_algo.h __partition),
 * the algorithm bvh_accel.cc:402 resolves to for unsigned int*. */
static uint32_t *partition_bidir(uint32_t *first, uint32_t *last, const ora_mesh *m, int axis, double pos) {
  for (;;) {
    for (;;) {
      if (first == last) return first;
      else if (sah_pred(m, *first, axis, pos)) ++first;
      else break;
    }
    --last;
    for (;;) {
      if (first == last) return first;
      else if (!sah_pred(m, *last, axis, pos)) --last;
      else break;
    }
    uint32_t t = *first; *first = *last; *last = t;
    ++first;
  }
}

/* BVHAccel::BuildTree, bvh_accel.cc:321-443 */
static size_t build_tree(build_ctx *c, uint32_t l, uint32_t r, int depth) {
  ora_bvh *b = c->bvh;
  size_t offset = b->nnodes;
  if (b->stats[0] < depth) b->stats[0] = depth;

  double mn[3], mx[3];
  compute_bbox(c, l, r, mn, mx);

  size_t n = (size_t)r - l;
  if (n < (size_t)c->min_leaf || depth >= c->max_depth) {
    ora_node leaf;
    memset(&leaf, 0, sizeof(leaf));
    memcpy(leaf.bmin, mn, sizeof(mn));
    memcpy(leaf.bmax, mx, sizeof(mx));
    leaf.flag = 1;
    leaf.axis = 0; /* uninitialised in the reference (bvh_accel.cc:343-360) */
    leaf.data[0] = (uint32_t)n;
    leaf.data[1] = l;
    push_node(b, &leaf);
    b->stats[1]++;
    return offset;
  }

  int axis = 0;
  double cut_pos[3] = {0.0, 0.0, 0.0};
  fill_bins(c, mn, mx, l, r);
  find_cut(c, mn, mx, n, cut_pos, &axis);

  /* the reference loop runs exactly once (axisTry < 1, bvh_accel.cc:389) */
  uint32_t *begin = &b->indices[l];
  uint32_t *end = begin + n;
  uint32_t *midp = partition_bidir(begin, end, c->mesh, axis, cut_pos[axis]);
  uint32_t mid = l + (uint32_t)(midp - begin);
  if (mid == l || mid == r) mid = l + (uint32_t)(n >> 1);

  ora_node node;
  memset(&node, 0, sizeof(node));
  node.axis = axis;
  node.flag = 0;
  push_node(b, &node);

  uint32_t lc = (uint32_t)build_tree(c, l, mid, depth + 1);
  uint32_t rc = (uint32_t)build_tree(c, mid, r, depth + 1);

  ora_node *me = &b->nodes[offset];
  me->data[0] = lc;
  me->data[1] = rc;
  memcpy(me->bmin, mn, sizeof(mn));
  memcpy(me->bmax, mx, sizeof(mx));
  b->stats[2]++;
  return offset;
}

/* BVHAccel::Build, bvh_accel.cc:445-482 */
ora_bvh *ora_bvh_build(const ora_mesh *mesh, double cost_taabb, int min_leaf, int max_depth, int bin_size) {
  ora_bvh *b = (ora_bvh *)calloc(1, sizeof(ora_bvh));
  size_t n = mesh->num_faces;
  b->nindices = n;
  b->indices = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) b->indices[i] = (uint32_t)i;
  if (n == 0) return b; /* the reference would throw from indices_.at(0); an empty tree is the sane result */
  build_ctx c;
  c.mesh = mesh;
  c.bvh = b;
  c.taabb = cost_taabb;
  c.min_leaf = min_leaf;
  c.max_depth = max_depth;
  c.nbins = bin_size;
  c.bins = (size_t *)malloc(sizeof(size_t) * 2 * 3 * (size_t)bin_size);
  build_tree(&c, 0, (uint32_t)n, 0);
  free(c.bins);
  return b;
}

ora_bvh *ora_bvh_from_arrays(const ora_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices) {
  ora_bvh *b = (ora_bvh *)calloc(1, sizeof(ora_bvh));
  b->nodes = (ora_node *)malloc(sizeof(ora_node) * (nnodes ? nnodes : 1));
  memcpy(b->nodes, nodes, sizeof(ora_node) * nnodes);
  b->nnodes = b->cap = nnodes;
  b->indices = (uint32_t *)malloc(sizeof(uint32_t) * (nindices ? nindices : 1));
  memcpy(b->indices, indices, sizeof(uint32_t) * nindices);
  b->nindices = nindices;
  return b;
}

void ora_bvh_free(ora_bvh *b) {
  if (!b) return;
  free(b->nodes);
  free(b->indices);
  free(b);
}
size_t ora_bvh_num_nodes(const ora_bvh *b) { return b->nnodes; }
size_t ora_bvh_num_indices(const ora_bvh *b) { return b->nindices; }
const ora_node *ora_bvh_nodes(const ora_bvh *b) { return b->nodes; }
const uint32_t *ora_bvh_indices(const ora_bvh *b) { return b->indices; }
void ora_bvh_stats(const ora_bvh *b, int out3[3]) { memcpy(out3, b->stats, sizeof(b->stats)); }

/* BVHAccel::Dump, bvh_accel.cc:484-513: u64 numNodes; BVHNode[]; u64 numIndices; u32[] */
int ora_bvh_dump(const ora_bvh *b, const char *path) {
  FILE *fp = fopen(path, "wb");
  if (!fp) return 0;
  unsigned long long nn = b->nnodes, ni = b->nindices;
  int ok = fwrite(&nn, sizeof(nn), 1, fp) == 1;
  ok = ok && fwrite(b->nodes, sizeof(ora_node), nn, fp) == nn;
  ok = ok && fwrite(&ni, sizeof(ni), 1, fp) == 1;
  ok = ok && fwrite(b->indices, sizeof(uint32_t), ni, fp) == ni;
  fclose(fp);
  return ok;
}

/* BVHAccel::Load, bvh_accel.cc:515-544 */
ora_bvh *ora_bvh_load(const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return NULL;
  unsigned long long nn = 0, ni = 0;
  ora_bvh *b = (ora_bvh *)calloc(1, sizeof(ora_bvh));
  int ok = fread(&nn, sizeof(nn), 1, fp) == 1 && nn > 0;
  if (ok) {
    b->nodes = (ora_node *)malloc(sizeof(ora_node) * nn);
    b->nnodes = b->cap = nn;
    ok = fread(b->nodes, sizeof(ora_node), nn, fp) == nn;
  }
  ok = ok && fread(&ni, sizeof(ni), 1, fp) == 1;
  if (ok) {
    b->indices = (uint32_t *)malloc(sizeof(uint32_t) * (ni ? ni : 1));
    b->nindices = ni;
    ok = fread(b->indices, sizeof(uint32_t), ni, fp) == ni;
  }
  fclose(fp);
  if (!ok) { ora_bvh_free(b); return NULL; }
  return b;
}

/* ======================================================================
 * Traversal -- bvh_accel.cc:546-844
 * ==================================================================== */

/* IntersectRayAABB, bvh_accel.cc:550-593: ternaries, not fmin/fmax. */
static inline int ray_aabb(double max_t, const double bmin[3], const double bmax[3], const double org[3],
                           const double inv[3], const int sign[3]) {
  const double min_x = sign[0] ? bmax[0] : bmin[0];
  const double min_y = sign[1] ? bmax[1] : bmin[1];
  const double min_z = sign[2] ? bmax[2] : bmin[2];
  const double max_x = sign[0] ? bmin[0] : bmax[0];
  const double max_y = sign[1] ? bmin[1] : bmax[1];
  const double max_z = sign[2] ? bmin[2] : bmax[2];

  const double tmin_x = (min_x - org[0]) * inv[0];
  const double tmax_x = (max_x - org[0]) * inv[0];
  const double tmin_y = (min_y - org[1]) * inv[1];
  const double tmax_y = (max_y - org[1]) * inv[1];
  double tmin = (tmin_x > tmin_y) ? tmin_x : tmin_y;
  double tmax = (tmax_x < tmax_y) ? tmax_x : tmax_y;
  const double tmin_z = (min_z - org[2]) * inv[2];
  const double tmax_z = (max_z - org[2]) * inv[2];
  tmin = (tmin > tmin_z) ? tmin : tmin_z;
  tmax = (tmax < tmax_z) ? tmax : tmax_z;
  return (tmax > 0.0) && (tmin <= tmax) && (tmin <= max_t);
}

/* TriangleIsect, bvh_accel.cc:595-638 */
static inline int tri_isect(double *t_io, double *u_out, double *v_out, v3 p0, v3 p1, v3 p2, v3 org, v3 dir) {
  v3 e1 = v3_sub(p1, p0);
  v3 e2 = v3_sub(p2, p0);
  v3 p = v3_cross(dir, e2);
  double det = v3_dot(e1, p);
  if (fabs(det) < ORA_KEPS) return 0;
  double inv_det = 1.0 / det;
  v3 s = v3_sub(org, p0);
  v3 q = v3_cross(s, e1);
  double u = v3_dot(s, p) * inv_det;
  double v = v3_dot(q, dir) * inv_det;
  double t = v3_dot(e2, q) * inv_det;
  if (u < 0.0 || u > 1.0) return 0;
  if (v < 0.0 || u + v > 1.0) return 0;
  if (t < 0.0 || t > *t_io) return 0;
  *t_io = t;
  *u_out = u;
  *v_out = v;
  return 1;
}

/* TestLeafNode, bvh_accel.cc:640-697 */
static int test_leaf(ora_isect *isect, const ora_node *node, const uint32_t *indices, const ora_mesh *mesh,
                     v3 org, v3 dir) {
  int hit = 0;
  uint32_t ntri = node->data[0], off = node->data[1];
  double t = isect->t;
  for (uint32_t i = 0; i < ntri; i++) {
    uint32_t f = indices[i + off];
    const uint32_t *fv = &mesh->faces[3 * (size_t)f];
    v3 a = v3_ptr(&mesh->vertices[3 * (size_t)fv[0]]);
    v3 b = v3_ptr(&mesh->vertices[3 * (size_t)fv[1]]);
    v3 c = v3_ptr(&mesh->vertices[3 * (size_t)fv[2]]);
    double u, v;
    if (tri_isect(&t, &u, &v, a, b, c, org, dir)) {
      isect->t = t;
      isect->u = u;
      isect->v = v;
      isect->face_id = f;
      isect->material_id = mesh->material_ids ? mesh->material_ids[f] : (uint32_t)-1;
      hit = 1;
    }
  }
  return hit;
}

/* BuildIntersection, bvh_accel.cc:699-769 */
static void build_isect(ora_isect *is, const ora_mesh *mesh, v3 org, v3 dir) {
  const uint32_t *fv = &mesh->faces[3 * (size_t)is->face_id];
  is->f0 = fv[0]; is->f1 = fv[1]; is->f2 = fv[2];
  v3 p0 = v3_ptr(&mesh->vertices[3 * (size_t)is->f0]);
  v3 p1 = v3_ptr(&mesh->vertices[3 * (size_t)is->f1]);
  v3 p2 = v3_ptr(&mesh->vertices[3 * (size_t)is->f2]);
  is->position[0] = org.x + is->t * dir.x;
  is->position[1] = org.y + is->t * dir.y;
  is->position[2] = org.z + is->t * dir.z;
  v3 n = v3_normalize(v3_cross(v3_sub(p1, p0), v3_sub(p2, p0)));
  is->geometric_normal[0] = n.x; is->geometric_normal[1] = n.y; is->geometric_normal[2] = n.z;
  if (mesh->fv_normals) {
    const double *N = &mesh->fv_normals[9 * (size_t)is->face_id];
    for (int k = 0; k < 3; k++)
      is->normal[k] = (1.0 - is->u - is->v) * N[k] + is->u * N[3 + k] + is->v * N[6 + k];
  } else {
    is->normal[0] = n.x; is->normal[1] = n.y; is->normal[2] = n.z;
  }
  if (mesh->fv_uvs) {
    const double *T = &mesh->fv_uvs[6 * (size_t)is->face_id];
    for (int k = 0; k < 2; k++)
      is->texcoord[k] = (1.0 - is->u - is->v) * T[k] + is->u * T[2 + k] + is->v * T[4 + k];
  }
}

#define ORA_STACK 512 /* kMaxStackDepth, bvh_accel.cc:548 */

/* BVHAccel::Traverse, bvh_accel.cc:773-844 */
int ora_traverse(const ora_bvh *b, const ora_mesh *mesh, const double org_[3], const double dir_[3],
                 ora_isect *isect, uint64_t counters[3]) {
  double hit_t = DBL_MAX;
  int sp = 0;
  int stack[ORA_STACK];
  stack[0] = 0;
  uint64_t n_node = 0, n_tri = 0, max_sp = 0;

  isect->t = hit_t;
  isect->u = 0.0;
  isect->v = 0.0;
  isect->face_id = (uint32_t)-1;

  if (b->nnodes == 0) return 0; /* guard: the reference would index an empty vector */

  int sign[3];
  double inv[3];
  for (int k = 0; k < 3; k++) {
    sign[k] = dir_[k] < 0.0 ? 1 : 0;
    inv[k] = 1.0 / dir_[k];
  }
  v3 org = v3_ptr(org_), dir = v3_ptr(dir_);

  while (sp >= 0) {
    const ora_node *node = &b->nodes[stack[sp]];
    sp--;
    n_node++;
    int hit = ray_aabb(hit_t, node->bmin, node->bmax, org_, inv, sign);
    if (node->flag == 0) {
      if (hit) {
        int near = sign[node->axis];
        int far = 1 - near;
        stack[++sp] = (int)node->data[far];
        stack[++sp] = (int)node->data[near];
        if ((uint64_t)(sp + 1) > max_sp) max_sp = (uint64_t)(sp + 1);
      }
    } else if (hit) {
      n_tri += node->data[0];
      if (test_leaf(isect, node, b->indices, mesh, org, dir)) hit_t = isect->t;
    }
  }
  if (counters) {
    counters[0] += n_node;
    counters[1] += n_tri;
    if (max_sp > counters[2]) counters[2] = max_sp;
  }
  if (isect->t < DBL_MAX) {
    build_isect(isect, mesh, org, dir);
    return 1;
  }
  return 0;
}

double ora_trace_batch(const ora_bvh *b, const ora_mesh *mesh, const double *rays, size_t n, ora_hit *hits,
                       ora_isect *isects, uint8_t *mask, uint64_t totals[3], int row, int nthreads) {
  if (row <= 0) row = 1920;
  long nrows = (long)((n + (size_t)row - 1) / (size_t)row);
  uint64_t tn = 0, tt = 0, ts = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  double t0 = ora_now();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tn, tt) reduction(max : ts)
  for (long y = 0; y < nrows; y++) {
    size_t lo = (size_t)y * row, hi = lo + row;
    if (hi > n) hi = n;
    for (size_t i = lo; i < hi; i++) {
      ora_isect is;
      memset(&is, 0, sizeof(is));
      uint64_t c[3] = {0, 0, 0};
      int hit = ora_traverse(b, mesh, &rays[6 * i], &rays[6 * i + 3], &is, c);
      tn += c[0];
      tt += c[1];
      if (c[2] > ts) ts = c[2];
      if (hits) {
        hits[i].t = is.t; hits[i].u = is.u; hits[i].v = is.v;
        hits[i].face_id = is.face_id; hits[i].material_id = is.material_id;
      }
      if (isects) isects[i] = is;
      if (mask) mask[i] = (uint8_t)hit;
    }
  }
  double t1 = ora_now();
  if (totals) { totals[0] += tn; totals[1] += tt; if (ts > totals[2]) totals[2] = ts; }
  return t1 - t0;
}

void ora_occluded_batch(const ora_bvh *b, const ora_mesh *mesh, const double *rays, const double *tmax, size_t n,
                        uint8_t *occluded, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024)
  for (long i = 0; i < (long)n; i++) {
    ora_isect is;
    memset(&is, 0, sizeof(is));
    int hit = ora_traverse(b, mesh, &rays[6 * i], &rays[6 * i + 3], &is, NULL);
    occluded[i] = (uint8_t)(hit && is.t < tmax[i]);
  }
}

/* ======================================================================
 * Camera -- camera.cc:12-240, matrix.cc:8-216, trackball.cc:272-292
 * ==================================================================== */
static inline double a_dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void a_cross(double c[3], const double a[3], const double b[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
/* vlength, camera.cc:22-28 == matrix.cc:18-24 */
static inline double a_length(const double v[3]) {
  double len2 = a_dot(v, v);
  if (fabs(len2) > 1.0e-30) return sqrt(len2);
  return 0.0;
}
/* matrix.cc:26-34: double length */
static void a_normalize_d(double v[3]) {
  double len = a_length(v);
  if (fabs(len) > 1.0e-30) {
    double inv = 1.0 / len;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
  }
}
/* camera.cc:30-38: the length is truncated to FLOAT before the reciprocal */
static void a_normalize_f(double v[3]) {
  float len = (float)a_length(v);
  if (fabsf(len) > 1.0e-30) {
    double inv = 1.0 / len;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
  }
}

/* Matrix::LookAt, matrix.cc:42-99 (the #else layout) */
static void mat_lookat(double m[4][4], const double eye[3], const double lookat[3], const double up[3]) {
  double u[3], v[3], look[3];
  for (int k = 0; k < 3; k++) look[k] = lookat[k] - eye[k];
  a_normalize_d(look);
  a_cross(u, look, up);
  a_normalize_d(u);
  a_cross(v, u, look);
  a_normalize_d(v);
  for (int k = 0; k < 3; k++) {
    m[0][k] = u[k];
    m[1][k] = v[k];
    m[2][k] = -look[k];
    m[3][k] = eye[k];
  }
  m[0][3] = 0.0; m[1][3] = 0.0; m[2][3] = 0.0; m[3][3] = 1.0;
}

/* Matrix::Inverse, matrix.cc:101-193: Cramer's rule on the transposed source,
 * cofactor pairs evaluated in the reference's order (every sum is left-assoc). */
static void mat_inverse(double m[4][4]) {
  double p[12], s[16], det;
  for (int i = 0; i < 4; i++) {
    s[i] = m[i][0]; s[i + 4] = m[i][1]; s[i + 8] = m[i][2]; s[i + 12] = m[i][3];
  }
  p[0] = s[10] * s[15]; p[1] = s[11] * s[14]; p[2] = s[9] * s[15]; p[3] = s[11] * s[13];
  p[4] = s[9] * s[14];  p[5] = s[10] * s[13]; p[6] = s[8] * s[15]; p[7] = s[11] * s[12];
  p[8] = s[8] * s[14];  p[9] = s[10] * s[12]; p[10] = s[8] * s[13]; p[11] = s[9] * s[12];

  m[0][0] = p[0] * s[5] + p[3] * s[6] + p[4] * s[7];
  m[0][0] -= p[1] * s[5] + p[2] * s[6] + p[5] * s[7];
  m[0][1] = p[1] * s[4] + p[6] * s[6] + p[9] * s[7];
  m[0][1] -= p[0] * s[4] + p[7] * s[6] + p[8] * s[7];
  m[0][2] = p[2] * s[4] + p[7] * s[5] + p[10] * s[7];
  m[0][2] -= p[3] * s[4] + p[6] * s[5] + p[11] * s[7];
  m[0][3] = p[5] * s[4] + p[8] * s[5] + p[11] * s[6];
  m[0][3] -= p[4] * s[4] + p[9] * s[5] + p[10] * s[6];
  m[1][0] = p[1] * s[1] + p[2] * s[2] + p[5] * s[3];
  m[1][0] -= p[0] * s[1] + p[3] * s[2] + p[4] * s[3];
  m[1][1] = p[0] * s[0] + p[7] * s[2] + p[8] * s[3];
  m[1][1] -= p[1] * s[0] + p[6] * s[2] + p[9] * s[3];
  m[1][2] = p[3] * s[0] + p[6] * s[1] + p[11] * s[3];
  m[1][2] -= p[2] * s[0] + p[7] * s[1] + p[10] * s[3];
  m[1][3] = p[4] * s[0] + p[9] * s[1] + p[10] * s[2];
  m[1][3] -= p[5] * s[0] + p[8] * s[1] + p[11] * s[2];

  p[0] = s[2] * s[7]; p[1] = s[3] * s[6]; p[2] = s[1] * s[7]; p[3] = s[3] * s[5];
  p[4] = s[1] * s[6]; p[5] = s[2] * s[5]; p[6] = s[0] * s[7]; p[7] = s[3] * s[4];
  p[8] = s[0] * s[6]; p[9] = s[2] * s[4]; p[10] = s[0] * s[5]; p[11] = s[1] * s[4];

  m[2][0] = p[0] * s[13] + p[3] * s[14] + p[4] * s[15];
  m[2][0] -= p[1] * s[13] + p[2] * s[14] + p[5] * s[15];
  m[2][1] = p[1] * s[12] + p[6] * s[14] + p[9] * s[15];
  m[2][1] -= p[0] * s[12] + p[7] * s[14] + p[8] * s[15];
  m[2][2] = p[2] * s[12] + p[7] * s[13] + p[10] * s[15];
  m[2][2] -= p[3] * s[12] + p[6] * s[13] + p[11] * s[15];
  m[2][3] = p[5] * s[12] + p[8] * s[13] + p[11] * s[14];
  m[2][3] -= p[4] * s[12] + p[9] * s[13] + p[10] * s[14];
  m[3][0] = p[2] * s[10] + p[5] * s[11] + p[1] * s[9];
  m[3][0] -= p[4] * s[11] + p[0] * s[9] + p[3] * s[10];
  m[3][1] = p[8] * s[11] + p[0] * s[8] + p[7] * s[10];
  m[3][1] -= p[6] * s[10] + p[9] * s[11] + p[1] * s[8];
  m[3][2] = p[6] * s[9] + p[11] * s[11] + p[3] * s[8];
  m[3][2] -= p[10] * s[11] + p[2] * s[8] + p[7] * s[9];
  m[3][3] = p[10] * s[10] + p[4] * s[8] + p[9] * s[9];
  m[3][3] -= p[8] * s[9] + p[11] * s[0] + p[5] * s[8]; /* sic: s[0], as matrix.cc:179 */

  det = s[0] * m[0][0] + s[1] * m[0][1] + s[2] * m[0][2] + s[3] * m[0][3];
  det = 1.0f / det;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) m[j][i] *= det;
}

/* Matrix::Mult, matrix.cc:195-204 */
static void mat_mult(double dst[4][4], double m0[4][4], double m1[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      dst[i][j] = 0;
      for (int k = 0; k < 4; ++k) dst[i][j] += m0[k][j] * m1[i][k];
    }
}

/* Matrix::MultV, matrix.cc:206-216 */
static void mat_multv(double dst[3], double m[4][4], const double v[3]) {
  dst[0] = m[0][0] * v[0] + m[1][0] * v[1] + m[2][0] * v[2] + m[3][0];
  dst[1] = m[0][1] * v[0] + m[1][1] * v[1] + m[2][1] * v[2] + m[3][1];
  dst[2] = m[0][2] * v[0] + m[1][2] * v[1] + m[2][2] * v[2] + m[3][2];
}

/* build_rotmatrix, trackball.cc:272-292 */
static void quat_to_matrix(double m[4][4], const double q[4]) {
  m[0][0] = 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]);
  m[0][1] = 2.0 * (q[0] * q[1] - q[2] * q[3]);
  m[0][2] = 2.0 * (q[2] * q[0] + q[1] * q[3]);
  m[0][3] = 0.0;
  m[1][0] = 2.0 * (q[0] * q[1] + q[2] * q[3]);
  m[1][1] = 1.0 - 2.0 * (q[2] * q[2] + q[0] * q[0]);
  m[1][2] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
  m[1][3] = 0.0;
  m[2][0] = 2.0 * (q[2] * q[0] - q[1] * q[3]);
  m[2][1] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
  m[2][2] = 1.0 - 2.0 * (q[1] * q[1] + q[0] * q[0]);
  m[2][3] = 0.0;
  m[3][0] = 0.0; m[3][1] = 0.0; m[3][2] = 0.0; m[3][3] = 1.0;
}

/* Camera::BuildCameraFrame, camera.cc:40-220 */
void ora_camera_frame(const double eye[3], const double lookat[3], const double up[3], double fov,
                      const double quat[4], int width, int height, double origin[3], double corner[3],
                      double u[3], double v[3]) {
  double r[4][4];
  quat_to_matrix(r, quat);

  double lo[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
  double dist = a_length(lo);
  double dir[3] = {0.0, 0.0, dist};

  mat_inverse(r);

  double rr[4][4], re[4][4];
  double zero[3] = {0.0, 0.0, 0.0};
  double local_up[3] = {0.0, 1.0, 0.0};
  mat_lookat(re, dir, zero, local_up);
  re[3][0] += eye[0];
  re[3][1] += eye[1];
  re[3][2] += (eye[2] - dist);
  mat_mult(rr, r, re);

  double eye1[3], lookat1[3];
  mat_multv(eye1, rr, zero);
  dir[2] = -dir[2];
  mat_multv(lookat1, rr, dir);
  /* up1 = M*up - eye1 is computed and then discarded: "Use original up vector" (camera.cc:142-144) */
  double up1[3] = {up[0], up[1], up[2]};

  double flen = (0.5f * (double)height / tanf(0.5f * (double)(fov * M_PI / 180.0f)));
  double look1[3] = {lookat1[0] - eye1[0], lookat1[1] - eye1[1], lookat1[2] - eye1[2]};
  a_cross(u, look1, up1);
  a_normalize_f(u);
  a_cross(v, look1, u);
  a_normalize_f(v);
  a_normalize_f(look1);
  for (int k = 0; k < 3; k++) look1[k] = flen * look1[k] + eye1[k];
  for (int k = 0; k < 3; k++) corner[k] = look1[k] - 0.5f * (width * u[k] + height * v[k]);
  for (int k = 0; k < 3; k++) origin[k] = eye1[k];
}

/* Camera::GenerateRay, camera.cc:222-240 */
void ora_generate_ray(const double origin[3], const double corner[3], const double du[3], const double dv[3],
                      double u, double v, double ray6[6]) {
  v3 d;
  d.x = (corner[0] + u * du[0] + v * dv[0]) - origin[0];
  d.y = (corner[1] + u * du[1] + v * dv[1]) - origin[1];
  d.z = (corner[2] + u * du[2] + v * dv[2]) - origin[2];
  d = v3_normalize(d);
  ray6[0] = origin[0]; ray6[1] = origin[1]; ray6[2] = origin[2];
  ray6[3] = d.x; ray6[4] = d.y; ray6[5] = d.z;
}

/* Camera::GenerateEnvRay, camera.cc:242-257, and GenerateStereoEnvRay, camera.cc:259-329 */
void ora_generate_env_ray(const double origin[3], int width, int height, double u, double v, int stereo,
                          double ray6[6]) {
  if (!stereo) {
    double theta = M_PI * (v / height);
    double phi = 2.0 * M_PI * (u / width);
    ray6[0] = origin[0]; ray6[1] = origin[1]; ray6[2] = origin[2];
    ray6[3] = sin(theta) * cos(phi);
    ray6[4] = cos(theta);
    ray6[5] = sin(theta) * sin(phi);
    return;
  }
  const int is_left_side = v < (height >> 1);
  const double focal_length = 4.0, r = 0.5;
  double theta = M_PI * fmod(2.0 * v / height, 1.0);
  double phi = 2.0 * M_PI * (u / width);
  v3 d0;
  d0.x = sin(theta) * cos(phi);
  d0.y = cos(theta);
  d0.z = sin(theta) * sin(phi);
  v3 par;
  if (is_left_side) { par.x = -d0.z; par.y = 0.0; par.z = d0.x; }
  else { par.x = d0.z; par.y = 0.0; par.z = -d0.x; }
  par = v3_normalize(par);
  par = v3_scale(par, r);
  ray6[0] = origin[0] + par.x; ray6[1] = origin[1] + par.y; ray6[2] = origin[2] + par.z;
  double psi = atan2(r, focal_length);
  if (is_left_side) psi = -psi;
  v3 d;
  d.x = d0.x * cos(psi) - d0.z * sin(psi);
  d.y = d0.y;
  d.z = d0.x * sin(psi) + d0.z * cos(psi);
  d = v3_normalize(d);
  ray6[3] = d.x; ray6[4] = d.y; ray6[5] = d.z;
}

void ora_generate_grid(const double origin[3], const double corner[3], const double du[3], const double dv[3],
                       int width, int height, double *rays) {
#pragma omp parallel for
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++)
      ora_generate_ray(origin, corner, du, dv, (double)x, (double)y, &rays[6 * ((size_t)y * width + x)]);
}

/* ======================================================================
 * Plane -- prim-plane.cc:8-44 (float vn, on_d, t)
 * ==================================================================== */
int ora_plane_intersect(const float abcd[4], const double org_[3], const double dir_[3], ora_isect *info) {
  v3 n = v3_make(abcd[0], abcd[1], abcd[2]);
  v3 v = v3_normalize(v3_ptr(dir_));
  v3 org = v3_ptr(org_);
  float vn = (float)v3_dot(v, n);
  if (fabsf(vn) > FLT_EPSILON * 1024.0f) {
    float on_d = (float)(v3_dot(org, n) + abcd[3]);
    float t = -on_d / vn;
    if ((t > 0) && (t < info->t)) {
      info->t = t;
      v3 tv = v3_scale(v, (double)t); /* real * real3 with float t promoted */
      info->position[0] = org.x + tv.x;
      info->position[1] = org.y + tv.y;
      info->position[2] = org.z + tv.z;
      n = v3_normalize(n);
      info->geometric_normal[0] = n.x; info->geometric_normal[1] = n.y; info->geometric_normal[2] = n.z;
      info->normal[0] = n.x; info->normal[1] = n.y; info->normal[2] = n.z;
      info->tangent[0] = 1.0; info->tangent[1] = 0.0; info->tangent[2] = 0.0;
      info->binormal[0] = 0.0; info->binormal[1] = 0.0; info->binormal[2] = -1.0;
      info->texcoord[0] = 0.0; info->texcoord[1] = 0.0;
      info->material_id = (uint32_t)-1;
      return 1;
    }
  }
  return 0;
}

/* render.cc:620-627 */
void ora_plane_from_bbox(const double bmin[3], const double bmax[3], float abcd[4]) {
  float zmin = (float)bmin[1];
  float zsize = (float)(bmax[1] - bmin[1]);
  abcd[0] = 0;
  abcd[1] = 1;
  abcd[2] = 0;
  abcd[3] = -(zmin - zsize * 0.0001f);
}

/* ======================================================================
 * RNG -- render.cc:116-168 (xorshift128)
 * ==================================================================== */
void ora_rng_seed_reference(ora_rng *r, int tid) {
  r->x = 123456789u + (uint32_t)tid;
  r->y = 362436069u;
  r->z = 521288629u;
  r->w = 88675123u;
}

static inline uint32_t mix32(uint32_t h) { /* murmur3 fmix32 */
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}

/* Counter-based seeding used by the CUDA path (mallie_b200/csrc/kernels: rng_seed_pixel):
 * the generator is still the reference's xorshift128; only the per-thread seed table
 * (gSeed[tid], render.cc:116-135) is replaced by a hash of (pixel index, pass). */
void ora_rng_seed_pixel(ora_rng *r, uint32_t pixel, uint32_t pass) {
  uint32_t k = mix32(pixel * 0x9e3779b9u + 0x7f4a7c15u) ^ mix32(pass * 0x85ebca6bu + 0x165667b1u);
  r->x = 123456789u ^ mix32(k + 1u);
  r->y = 362436069u ^ mix32(k + 2u);
  r->z = 521288629u ^ mix32(k + 3u);
  r->w = 88675123u ^ mix32(k + 4u);
  if ((r->x | r->y | r->z | r->w) == 0u) r->w = 88675123u;
}

double ora_randomreal(ora_rng *r) {
  uint32_t t = r->x ^ (r->x << 11);
  r->x = r->y;
  r->y = r->z;
  r->z = r->w;
  r->w = (r->w ^ (r->w >> 19)) ^ (t ^ (t >> 8));
  return r->w * (1.0 / 4294967296.0);
}

/* ======================================================================
 * Shading -- render.cc:271-339, 381-456
 * ==================================================================== */
/* GenerateBasis, render.cc:271-318 (fabsf: float magnitude picks the minor axis) */
static void gen_basis(v3 *tangent, v3 *binormal, v3 normal) {
  int index = -1;
  double minval = 1.0e+6;
  for (int i = 0; i < 3; i++) {
    double val = fabsf((float)v3_get(&normal, i));
    if (val < minval) { minval = val; index = i; }
  }
  if (index == 0) *tangent = v3_make(0.0, -normal.z, normal.y);
  else if (index == 1) *tangent = v3_make(-normal.z, 0.0, normal.x);
  else *tangent = v3_make(-normal.y, normal.x, 0.0);
  *tangent = v3_normalize(*tangent);
  *binormal = v3_normalize(v3_cross(*tangent, normal));
}

/* SampleDiffuseIS, render.cc:320-339 */
static v3 sample_diffuse(ora_rng *rng, v3 normal) {
  v3 tangent, binormal;
  gen_basis(&tangent, &binormal, normal);
  double theta = acos(sqrt(1.0 - ora_randomreal(rng)));
  double phi = 2.0 * M_PI * ora_randomreal(rng);
  double cos_theta = cos(theta);
  v3 T = v3_scale(v3_scale(tangent, cos(phi)), sin(theta));
  v3 B = v3_scale(v3_scale(binormal, sin(phi)), sin(theta));
  v3 N = v3_scale(normal, cos_theta);
  return v3_add(v3_add(T, B), N);
}

static const double kRenderEPS = 1.0e-3; /* render.cc:51 */
static const double kFar = 1.0e+30;      /* render.cc:50 */

/* PathTrace, render.cc:381-456.  With skip_zombies the
 * post-escape segments -- which can never hit -- are not traced; their
 * contribution throughput*0.5/pathLength is accumulated in the same order. */
static v3 path_trace(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, ora_rng *rng, int px,
                     int py, uint64_t cnt[7]) {
  float ju = (float)(ora_randomreal(rng) - 0.5);
  float jv = (float)(ora_randomreal(rng) - 0.5);
  double ray[6];
  /* PathTraceEnv (render.cc:518-590) is PathTrace with a panorama camera, without the plane and without
   * the material attenuation (its `throughput` is never used; the miss term is kd / pathLength, kd = 0.5) */
  const int env = p->shader == 2;
  if (p->camera_mode == 0) ora_generate_ray(p->origin, p->corner, p->du, p->dv, (double)(px + ju), (double)(py + jv), ray);
  else ora_generate_env_ray(p->origin, p->width, p->height, (double)(px + ju), (double)(py + jv), p->camera_mode == 2, ray);

  ora_isect is;
  memset(&is, 0, sizeof(is));
  is.t = kFar;
  v3 thr = v3_make(1.0, 1.0, 1.0), rad = v3_make(0.0, 0.0, 0.0);
  int escaped = 0;
  for (unsigned int len = 1;; ++len) {
    int hit = 0;
    if (escaped && p->skip_zombies) {
      cnt[1]++;
    } else {
      if (escaped) cnt[1]++;
      cnt[0]++;
      {
        uint64_t tc[3] = {0, 0, 0};
        hit = ora_traverse(b, mesh, ray, ray + 3, &is, tc);
        cnt[3] += tc[0];
        cnt[4] += tc[1];
      }
      if (p->use_plane && !env) hit |= ora_plane_intersect(p->plane, ray, ray + 3, &is);
    }
    if (!hit) {
      if (len < 2) break; /* kMinPathLength */
      double l = (double)len;
      rad.x += thr.x * 0.5 / l; rad.y += thr.y * 0.5 / l; rad.z += thr.z * 0.5 / l;
      escaped = 1;
    }
    if (len >= (unsigned int)p->max_path_length) break;

    if (escaped && p->skip_zombies) {
      /* the three RNG draws are unobservable with per-pixel seeding; the sequential reference
       * stream (rng_mode 0) must still consume them to stay aligned for the following pixels */
      if (p->rng_mode == 0) { (void)ora_randomreal(rng); (void)ora_randomreal(rng); (void)ora_randomreal(rng); }
      if (!env && is.material_id != (uint32_t)-1) { thr.x *= 0.5; thr.y *= 0.5; thr.z *= 0.5; }
      continue;
    }
    v3 org = v3_ptr(ray), dir = v3_ptr(ray + 3);
    v3 hit_p = v3_add(org, v3_scale(dir, is.t));
    (void)ora_randomreal(rng); /* drawn, unused: render.cc:430 */
    v3 n = v3_ptr(is.normal);
    double ndoti = v3_dot(n, v3_neg(dir));
    if (ndoti < 0.0) n = v3_neg(n);
    v3 nd = sample_diffuse(rng, n);
    if (!env && is.material_id != (uint32_t)-1) { /* Scene::GetMaterial -> default diffuse 0.5, scene.h:58-65 */
      thr.x *= 0.5; thr.y *= 0.5; thr.z *= 0.5;
    }
    v3 no = v3_add(hit_p, v3_scale(nd, kRenderEPS));
    ray[0] = no.x; ray[1] = no.y; ray[2] = no.z;
    ray[3] = nd.x; ray[4] = nd.y; ray[5] = nd.z;
    is.t = kFar;
  }
  return rad;
}

/* Primary + shadow shader: the next-event-estimation block the reference leaves
 * empty (render.cc:425-426), defined in DESIGN.md §"primary+shadow".  One closest-hit
 * primary ray; on a hit, one occlusion ray towards the point light. */
static v3 primary_shadow(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, ora_rng *rng, int px,
                         int py, uint64_t cnt[7], double *primary_out, double *shadow_out) {
  float ju = (float)(ora_randomreal(rng) - 0.5);
  float jv = (float)(ora_randomreal(rng) - 0.5);
  double ray[6];
  ora_generate_ray(p->origin, p->corner, p->du, p->dv, (double)(px + ju), (double)(py + jv), ray);
  ora_isect is;
  memset(&is, 0, sizeof(is));
  is.t = kFar;
  cnt[0]++;
  uint64_t tc[3] = {0, 0, 0};
  if (primary_out) memcpy(primary_out, ray, sizeof(ray));
  int hit = ora_traverse(b, mesh, ray, ray + 3, &is, tc);
  if (p->use_plane) hit |= ora_plane_intersect(p->plane, ray, ray + 3, &is);
  if (!hit) {
    cnt[3] += tc[0];
    cnt[4] += tc[1];
    return v3_make(0.0, 0.0, 0.0);
  }
  v3 org = v3_ptr(ray), dir = v3_ptr(ray + 3);
  v3 hit_p = v3_add(org, v3_scale(dir, is.t));
  v3 n = v3_ptr(is.normal);
  if (v3_dot(n, v3_neg(dir)) < 0.0) n = v3_neg(n);
  v3 l = v3_sub(v3_ptr(p->light), hit_p);
  double dist = sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
  v3 ld = v3_normalize(l);
  v3 so = v3_add(hit_p, v3_scale(ld, kRenderEPS));
  double sray[6] = {so.x, so.y, so.z, ld.x, ld.y, ld.z};
  double tmax = dist - kRenderEPS;
  ora_isect sh;
  memset(&sh, 0, sizeof(sh));
  cnt[2]++;
  if (shadow_out) { memcpy(shadow_out, sray, sizeof(sray)); shadow_out[6] = tmax; }
  const uint64_t pn = tc[0], pt = tc[1]; /* the camera ray's share */
  int occ = ora_traverse(b, mesh, sray, sray + 3, &sh, tc) && sh.t < tmax;
  cnt[3] += tc[0];
  cnt[4] += tc[1];
  cnt[5] += tc[0] - pn; /* the shadow ray's own node / triangle counts */
  cnt[6] += tc[1] - pt;
  double ndotl = v3_dot(n, ld);
  if (occ || !(ndotl > 0.0)) return v3_make(0.0, 0.0, 0.0);
  double kd = (is.material_id != (uint32_t)-1) ? 0.5 : 1.0;
  double c = kd * ndotl;
  return v3_make(c, c, c);
}

/* One pixel sample with either shader. */
static v3 shade_pixel(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, ora_rng *rng, int x, int y,
                      uint64_t c[7], double *primary_out, double *shadow_out) {
  if (p->shader == 0 || p->shader == 2) return path_trace(b, mesh, p, rng, x, y, c);
  return primary_shadow(b, mesh, p, rng, x, y, c, primary_out, shadow_out);
}

/* Render, render.cc:593-708 (OpenMP scanline loop :657-698; step == 1).
 * primary_rays_out (nullable, shader 1): [6*W*H] the jittered camera rays; shadow_rays_out (nullable):
 * [7*W*H] org, dir, tmax of each shadow ray (NaN-filled where the primary ray missed). */
void ora_render_pass_ex(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, int x0, int y0, int x1,
                        int y1, float *image, int *count, uint64_t ray_counts[7], int nthreads,
                        double *primary_rays_out, double *shadow_rays_out) {
  const int W = p->width;
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
  if (shadow_rays_out)
    for (size_t i = 0; i < (size_t)7 * W * p->height; i++) shadow_rays_out[i] = NAN;
  if (p->rng_mode == 0) {
    /* sequential stream of OpenMP thread 0: identical to the reference with OMP_NUM_THREADS=1.  The
     * reference keeps the stream across Render() calls (gSeed is global); only the first pass is
     * reproducible here. */
    ora_rng rng;
    ora_rng_seed_reference(&rng, 0);
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) {
        uint64_t c[7] = {0, 0, 0, 0, 0, 0, 0};
        size_t pix = (size_t)y * W + x;
        v3 r = shade_pixel(b, mesh, p, &rng, x, y, c, primary_rays_out ? primary_rays_out + 6 * pix : NULL,
                           shadow_rays_out ? shadow_rays_out + 7 * pix : NULL);
        image[3 * pix + 0] = (float)r.x;
        image[3 * pix + 1] = (float)r.y;
        image[3 * pix + 2] = (float)r.z;
        count[pix]++;
        c0 += c[0]; c1 += c[1]; c2 += c[2]; c3 += c[3]; c4 += c[4]; c5 += c[5]; c6 += c[6];
      }
  } else {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1, c2, c3, c4, c5, c6)
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) {
        ora_rng rng;
        size_t pix = (size_t)y * W + x;
        ora_rng_seed_pixel(&rng, (uint32_t)pix, p->pass);
        uint64_t c[7] = {0, 0, 0, 0, 0, 0, 0};
        v3 r = shade_pixel(b, mesh, p, &rng, x, y, c, primary_rays_out ? primary_rays_out + 6 * pix : NULL,
                           shadow_rays_out ? shadow_rays_out + 7 * pix : NULL);
        image[3 * pix + 0] = (float)r.x;
        image[3 * pix + 1] = (float)r.y;
        image[3 * pix + 2] = (float)r.z;
        count[pix]++;
        c0 += c[0]; c1 += c[1]; c2 += c[2]; c3 += c[3]; c4 += c[4]; c5 += c[5]; c6 += c[6];
      }
  }
  if (ray_counts) {
    ray_counts[0] += c0; ray_counts[1] += c1; ray_counts[2] += c2; ray_counts[3] += c3; ray_counts[4] += c4;
    ray_counts[5] += c5; ray_counts[6] += c6;
  }
}

void ora_render_pass(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p, int x0, int y0, int x1,
                     int y1, float *image, int *count, uint64_t ray_counts[7], int nthreads) {
  ora_render_pass_ex(b, mesh, p, x0, y0, x1, y1, image, count, ray_counts, nthreads, NULL, NULL);
}

/* RenderPanoramic, render.cc:710-763 */
void ora_render_panoramic(const ora_bvh *b, const ora_mesh *mesh, const ora_render_params *p0, float *image,
                          int *count, int nthreads) {
  ora_render_params p = *p0;
  p.shader = 2;
  p.use_plane = 0;
  const int W = p.width, H = p.height;
  memset(image, 0, sizeof(float) * (size_t)W * H * 3);
  if (p.rng_mode == 0) {
    ora_rng rng;
    ora_rng_seed_reference(&rng, 0);
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++)
        for (int i = 0; i < 10; i++) {
          uint64_t c[7] = {0, 0, 0, 0, 0, 0, 0};
          size_t pix = (size_t)y * W + x;
          v3 r = path_trace(b, mesh, &p, &rng, x, y, c);
          image[3 * pix + 0] += r.x;
          image[3 * pix + 1] += r.y;
          image[3 * pix + 2] += r.z;
          count[pix]++;
        }
    return;
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int i = 0; i < 10; i++) {
        ora_rng rng;
        size_t pix = (size_t)y * W + x;
        ora_rng_seed_pixel(&rng, (uint32_t)pix, p0->pass + (uint32_t)i);
        uint64_t c[7] = {0, 0, 0, 0, 0, 0, 0};
        v3 r = path_trace(b, mesh, &p, &rng, x, y, c);
        image[3 * pix + 0] += r.x;
        image[3 * pix + 1] += r.y;
        image[3 * pix + 2] += r.z;
        count[pix]++;
      }
}

/* ======================================================================
 * FNV-1a-64 (SURVEY App. B hashing convention)
 * ==================================================================== */
/* ======================================================================
 * Output resolve -- main_console.cc:25-43, main_sdl.cc:156-165,420-477
 * ==================================================================== */
/* fclamp, main_console.cc:25-32: `int i = x * 255.5;` is float * double -> double -> int (cvttsd2si) */
static unsigned char fclamp_console(float x) {
  int i = x * 255.5;
  if (i < 0) return 0;
  if (i > 255) return 255;
  return (unsigned char)i;
}
void ora_hdr_to_ldr(const float *in, const int *in_count, int width, int height, unsigned char *out) {
  for (long i = 0; i < (long)width * height * 3; i++) out[i] = fclamp_console(in[i] / in_count[i / 3]);
}
/* fclamp, main_sdl.cc:156-165 */
static unsigned char fclamp_display(float x) {
  float gamma = 2.2f;
  int i = powf(x, 1.0f / gamma) * 255.5;
  if (i < 0) return 0;
  if (i > 255) return 255;
  return (unsigned char)i;
}
void ora_display_bgra(const float *in, const int *counts, int width, int height, unsigned char *out) {
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const long p = (long)y * width + x;
      float scale = 1.0f / (float)counts[p];
      out[4 * p + 2] = fclamp_display(scale * in[3 * p + 0]);
      out[4 * p + 1] = fclamp_display(scale * in[3 * p + 1]);
      out[4 * p + 0] = fclamp_display(scale * in[3 * p + 2]);
      out[4 * p + 3] = 255;
    }
}

uint64_t ora_fnv1a64(const void *data, size_t nbytes, uint64_t seed) {
  const unsigned char *p = (const unsigned char *)data;
  uint64_t h = seed ? seed : 14695981039346656037ULL;
  for (size_t i = 0; i < nbytes; i++) {
    h ^= p[i];
    h *= 1099511628211ULL;
  }
  return h;
}

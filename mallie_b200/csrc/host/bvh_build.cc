// bvh_build.cc -- host-side binned-SAH BVH builder.
//
// Produces, bit for bit, the tree BVHAccel::Build makes in the reference
// (bvh_accel.cc:36-482; algorithm restated in SURVEY.md App. A.1): same
// pre-order node array, same permutation of the triangle index array, same
// (double) node bounds.  That identity is what makes faceID tie-breaks on the
// GPU agree with the reference, so it is a correctness requirement, not a
// nicety.
//
// It is not a transcription.  Differences that do not change the result:
//   * per-triangle bounds and centroid sums are computed once up front instead
//     of being re-gathered through faces[]/vertices[] at every tree level;
//   * node bounds use  min_i(v_i) - kEPS  ==  min_i(v_i - kEPS)  (rounding is
//     monotonic), so they come from the cached triangle bounds;
//   * independent subtrees are built by OpenMP tasks into private node vectors
//     and spliced into pre-order afterwards (child indices are rebased), which
//     cuts the 10 M-triangle build from ~17 s to a few seconds;
//   * the partition is libstdc++'s bidirectional std::partition algorithm
//     (what std::partition(unsigned*, ...) resolves to, bvh_accel.cc:402),
//     written out so the in-leaf triangle order does not depend on the STL.
#include "bvh_build.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace mb200 {

namespace {

const double kBoundsPad = DBL_EPSILON * 1024.0; // bvh_accel.cc:283

struct TriCache {
  // SoA, one entry per triangle id.
  std::vector<double> lo[3], hi[3], csum[3];
};

struct Builder {
  const TriCache *tc;
  uint32_t *indices;
  mb200_build_options opt;
  const double *vertices; // for the box of an empty range only
  const uint32_t *faces;
  size_t nfaces;
};

struct SubTree {
  std::vector<mb200_bvh_node> nodes; // child indices local to this vector
  int max_depth = 0, leaves = 0, branches = 0;
};

inline double half_area2(const double lo[3], const double hi[3]) {
  // CalculateSurfaceArea (bvh_accel.cc:50-53)
  double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return 2.0 * (dx * dy + dy * dz + dz * dx);
}

void range_bounds(const Builder &b, uint32_t l, uint32_t r, double lo[3], double hi[3]) {
  const TriCache &tc = *b.tc;
  if (l == r) {
    // An empty range (the left child of the object-median fallback on a single triangle, reachable only with
    // minLeafPrimitives <= 1): ComputeBoundingBox seeds the box with the first vertex of the triangle at
    // indices[leftIndex] before its loop (bvh_accel.cc:291-298) and the loop then adds nothing.
    const uint32_t t = b.indices[l < b.nfaces ? l : b.nfaces - 1];
    const double *p0 = b.vertices + 3 * (size_t)b.faces[3 * (size_t)t];
    for (int a = 0; a < 3; a++) lo[a] = p0[a] - kBoundsPad, hi[a] = p0[a] + kBoundsPad;
    return;
  }
  for (int a = 0; a < 3; a++) {
    double mn = tc.lo[a][b.indices[l]], mx = tc.hi[a][b.indices[l]];
    for (uint32_t i = l + 1; i < r; i++) {
      uint32_t t = b.indices[i];
      double x = tc.lo[a][t], y = tc.hi[a][t];
      if (x < mn) mn = x;
      if (y > mx) mx = y;
    }
    lo[a] = mn - kBoundsPad;
    hi[a] = mx + kBoundsPad;
  }
}

// Chooses the split plane: 64-bin histograms of triangle-bound minima/maxima per
// axis (ContributeBinBuffer, bvh_accel.cc:82-142) swept for the cheapest SAH
// plane (FindCutFromBinBuffer, bvh_accel.cc:156-255).
void choose_split(const Builder &b, uint32_t l, uint32_t r, const double lo[3], const double hi[3], int *axis_out,
                  double *pos_out) {
  const int nb = b.opt.bin_size;
  const double nbins = (double)nb;
  const TriCache &tc = *b.tc;
  std::vector<size_t> hist((size_t)2 * 3 * nb, 0);
  size_t *hmin = hist.data(), *hmax = hist.data() + (size_t)3 * nb;

  double scale[3], extent[3], step[3];
  for (int a = 0; a < 3; a++) {
    extent[a] = hi[a] - lo[a];
    scale[a] = (extent[a] > kBoundsPad) ? nbins / extent[a] : 0.0;
    step[a] = extent[a] * (1.0 / nb);
  }
  for (int a = 0; a < 3; a++) {
    const double *tlo = tc.lo[a].data(), *thi = tc.hi[a].data();
    for (uint32_t i = l; i < r; i++) {
      uint32_t t = b.indices[i];
      size_t qlo = (unsigned int)floor((tlo[t] - lo[a]) * scale[a]);
      size_t qhi = (unsigned int)floor((thi[t] - lo[a]) * scale[a]);
      if ((double)qlo >= nbins) qlo = (size_t)nb - 1;
      if ((double)qhi >= nbins) qhi = (size_t)nb - 1;
      hmin[a * nb + qlo]++;
      hmax[a * nb + qhi]++;
    }
  }

  const size_t n = (size_t)r - l;
  const double t_box = b.opt.cost_taabb, t_tri = 1.0 - b.opt.cost_taabb;
  const double whole = half_area2(lo, hi);
  const double inv_whole = (whole > kBoundsPad) ? 1.0 / whole : 0.0;
  double best_cost[3], best_pos[3];
  for (int a = 0; a < 3; a++) {
    best_pos[a] = lo[a] + 0.5 * step[a];
    best_cost[a] = DBL_MAX;
    size_t nl = 0, nr = n;
    double llo[3] = {lo[0], lo[1], lo[2]}, lhi[3] = {hi[0], hi[1], hi[2]};
    double rlo[3] = {lo[0], lo[1], lo[2]}, rhi[3] = {hi[0], hi[1], hi[2]};
    for (int i = 0; i < nb - 1; ++i) {
      nl += hmin[a * nb + i];
      nr -= hmax[a * nb + i];
      double pos = lo[a] + (i + 0.5) * step[a];
      lhi[a] = pos;
      rlo[a] = pos;
      double al = half_area2(llo, lhi), ar = half_area2(rlo, rhi);
      // SAH (bvh_accel.cc:144-154), evaluation order preserved
      double cost = 2.0f * t_box + (al * inv_whole) * (double)(nl)*t_tri + (ar * inv_whole) * (double)(nr)*t_tri;
      if (cost < best_cost[a]) {
        best_cost[a] = cost;
        best_pos[a] = pos;
      }
    }
  }
  int axis = 0;
  double c = best_cost[0];
  if (c > best_cost[1]) axis = 1, c = best_cost[1];
  if (c > best_cost[2]) axis = 2, c = best_cost[2];
  *axis_out = axis;
  *pos_out = best_pos[axis];
}

// libstdc++ __partition (bidirectional iterators); predicate SAHPred (bvh_accel.cc:257-277):
// triangle goes left iff (p0[a] + p1[a] + p2[a]) < pos * 3.0.
uint32_t split_range(const Builder &b, uint32_t l, uint32_t r, int axis, double pos) {
  const double *cs = b.tc->csum[axis].data();
  const double thresh = pos * 3.0;
  uint32_t *first = b.indices + l, *last = b.indices + r;
  for (;;) {
    for (;;) {
      if (first == last) return (uint32_t)(first - b.indices);
      if (cs[*first] < thresh) ++first;
      else break;
    }
    --last;
    for (;;) {
      if (first == last) return (uint32_t)(first - b.indices);
      if (!(cs[*last] < thresh)) --last;
      else break;
    }
    std::swap(*first, *last);
    ++first;
  }
}

const uint32_t kTaskCutoff = 1u << 15; // ranges smaller than this are built inline

void build_range(const Builder &b, uint32_t l, uint32_t r, int depth, SubTree &out);

void splice(SubTree &dst, size_t self, SubTree &left, SubTree &right) {
  const uint32_t lbase = (uint32_t)dst.nodes.size();
  for (mb200_bvh_node nd : left.nodes) {
    if (nd.flag == 0) nd.data[0] += lbase, nd.data[1] += lbase;
    dst.nodes.push_back(nd);
  }
  const uint32_t rbase = (uint32_t)dst.nodes.size();
  for (mb200_bvh_node nd : right.nodes) {
    if (nd.flag == 0) nd.data[0] += rbase, nd.data[1] += rbase;
    dst.nodes.push_back(nd);
  }
  dst.nodes[self].data[0] = lbase;
  dst.nodes[self].data[1] = rbase;
  dst.max_depth = std::max(dst.max_depth, std::max(left.max_depth, right.max_depth));
  dst.leaves += left.leaves + right.leaves;
  dst.branches += left.branches + right.branches;
}

// BuildTree (bvh_accel.cc:321-443).  Appends the subtree of [l, r) to out.nodes in pre-order.
void build_range(const Builder &b, uint32_t l, uint32_t r, int depth, SubTree &out) {
  const size_t self = out.nodes.size();
  if (out.max_depth < depth) out.max_depth = depth;

  mb200_bvh_node nd;
  memset(&nd, 0, sizeof(nd));
  range_bounds(b, l, r, nd.bmin, nd.bmax);

  const size_t n = (size_t)r - l;
  if (n < (size_t)b.opt.min_leaf_primitives || depth >= b.opt.max_tree_depth) {
    nd.flag = 1;
    nd.axis = 0; // never read; the reference leaves it uninitialised
    nd.data[0] = (uint32_t)n;
    nd.data[1] = l;
    out.nodes.push_back(nd);
    out.leaves++;
    return;
  }

  int axis;
  double pos;
  choose_split(b, l, r, nd.bmin, nd.bmax, &axis, &pos);
  uint32_t mid = split_range(b, l, r, axis, pos);
  if (mid == l || mid == r) mid = l + (uint32_t)(n >> 1); // object-median fallback, array left as partitioned

  nd.flag = 0;
  nd.axis = axis;
  out.nodes.push_back(nd);
  out.branches++;

  if (n >= kTaskCutoff) {
    SubTree left, right;
#pragma omp task shared(left) firstprivate(l, mid, depth)
    build_range(b, l, mid, depth + 1, left);
#pragma omp task shared(right) firstprivate(mid, r, depth)
    build_range(b, mid, r, depth + 1, right);
#pragma omp taskwait
    splice(out, self, left, right);
  } else {
    out.nodes[self].data[0] = (uint32_t)out.nodes.size();
    build_range(b, l, mid, depth + 1, out);
    out.nodes[self].data[1] = (uint32_t)out.nodes.size();
    build_range(b, mid, r, depth + 1, out);
  }
}

} // namespace

bool build_bvh(HostBVH &out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
               const mb200_build_options &opt, std::string *err) {
  out.nodes.clear();
  out.indices.clear();
  out.stats = mb200_build_stats{0, 0, 0};
  if (opt.bin_size <= 1 || opt.bin_size > 65536) {
    if (err) *err = "bin_size must be in (1, 65536]";
    return false;
  }
  // minLeafPrimitives <= 0 would split empty ranges into two empty ranges down to maxTreeDepth (2^depth nodes; the
  // reference does exactly that and runs out of memory); minLeafPrimitives == 1 is accepted and reproduces the
  // reference's depth-maxTreeDepth chains of empty left children under every single-triangle range.
  if (opt.min_leaf_primitives < 1) {
    if (err) *err = "min_leaf_primitives must be >= 1";
    return false;
  }
  if (opt.max_tree_depth < 0) {
    if (err) *err = "max_tree_depth must be >= 0";
    return false;
  }
  if (nfaces > 0xFFFFFFF0ull) {
    if (err) *err = "too many triangles (index array is 32-bit, bvh_accel.h:85)";
    return false;
  }
  for (size_t i = 0; i < 3 * nfaces; i++)
    if (faces[i] >= nverts) {
      if (err) *err = "face references a vertex out of range";
      return false;
    }
  out.indices.resize(nfaces);
  for (size_t i = 0; i < nfaces; i++) out.indices[i] = (uint32_t)i;
  if (nfaces == 0) return true; // empty tree: every ray misses

  TriCache tc;
  for (int a = 0; a < 3; a++) {
    tc.lo[a].resize(nfaces);
    tc.hi[a].resize(nfaces);
    tc.csum[a].resize(nfaces);
  }
#pragma omp parallel for schedule(static)
  for (long t = 0; t < (long)nfaces; t++) {
    const double *p0 = vertices + 3 * (size_t)faces[3 * t + 0];
    const double *p1 = vertices + 3 * (size_t)faces[3 * t + 1];
    const double *p2 = vertices + 3 * (size_t)faces[3 * t + 2];
    for (int a = 0; a < 3; a++) {
      double mn = p0[a], mx = p0[a];
      if (p1[a] < mn) mn = p1[a];
      if (mx < p1[a]) mx = p1[a];
      if (p2[a] < mn) mn = p2[a];
      if (mx < p2[a]) mx = p2[a];
      tc.lo[a][t] = mn;
      tc.hi[a][t] = mx;
      tc.csum[a][t] = p0[a] + p1[a] + p2[a];
    }
  }

  Builder b;
  b.tc = &tc;
  b.indices = out.indices.data();
  b.opt = opt;
  b.vertices = vertices, b.faces = faces, b.nfaces = nfaces;
  SubTree root;
  root.nodes.reserve(nfaces / 4 + 16);
#pragma omp parallel
#pragma omp single nowait
  build_range(b, 0, (uint32_t)nfaces, 0, root);

  out.nodes.swap(root.nodes);
  out.stats.max_tree_depth = root.max_depth;
  out.stats.num_leaf_nodes = root.leaves;
  out.stats.num_branch_nodes = root.branches;
  return true;
}

// BVHAccel::Dump (bvh_accel.cc:484-513): u64 numNodes, BVHNode[numNodes], u64 numIndices, u32[numIndices].
bool dump_bvh(const HostBVH &bvh, const char *path, std::string *err) {
  FILE *fp = fopen(path, "wb");
  if (!fp) {
    if (err) *err = std::string("cannot write ") + path;
    return false;
  }
  unsigned long long nn = bvh.nodes.size(), ni = bvh.indices.size();
  bool ok = fwrite(&nn, sizeof(nn), 1, fp) == 1;
  ok = ok && (nn == 0 || fwrite(bvh.nodes.data(), sizeof(mb200_bvh_node), nn, fp) == nn);
  ok = ok && fwrite(&ni, sizeof(ni), 1, fp) == 1;
  ok = ok && (ni == 0 || fwrite(bvh.indices.data(), sizeof(uint32_t), ni, fp) == ni);
  fclose(fp);
  if (!ok && err) *err = std::string("short write to ") + path;
  return ok;
}

// BVHAccel::Load (bvh_accel.cc:515-544).  Statistics are recomputed from the tree.
bool load_bvh(HostBVH &out, const char *path, std::string *err) {
  FILE *fp = fopen(path, "rb");
  if (!fp) {
    if (err) *err = std::string("cannot open ") + path;
    return false;
  }
  unsigned long long nn = 0, ni = 0;
  bool ok = fread(&nn, sizeof(nn), 1, fp) == 1 && nn > 0 && nn < (1ull << 32);
  if (ok) {
    out.nodes.resize(nn);
    ok = fread(out.nodes.data(), sizeof(mb200_bvh_node), nn, fp) == nn;
  }
  ok = ok && fread(&ni, sizeof(ni), 1, fp) == 1 && ni < (1ull << 32);
  if (ok) {
    out.indices.resize(ni);
    ok = ni == 0 || fread(out.indices.data(), sizeof(uint32_t), ni, fp) == ni;
  }
  fclose(fp);
  if (!ok) {
    if (err) *err = std::string("malformed BVH file ") + path;
    out.nodes.clear();
    out.indices.clear();
    return false;
  }
  out.stats = mb200_build_stats{0, 0, 0};
  for (const mb200_bvh_node &nd : out.nodes) (nd.flag ? out.stats.num_leaf_nodes : out.stats.num_branch_nodes)++;
  return true;
}

} // namespace mb200

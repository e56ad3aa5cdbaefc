// layout.h -- how a Mallie scene is laid out in HBM for the sm_100a kernels.
//
// The reference walks 64-byte BVHNode records that each carry their OWN box
// (bvh_accel.h:10-29) and fetches triangles through two indirections
// (indices_[i] -> faces[3f+k] -> vertices[3v+c], bvh_accel.cc:661-678).  On the
// device the same tree is stored "children-in-parent":
//
//   PairNode (128 B, 128-B aligned, one per reference BRANCH node)
//     box[c][0..2] = child c bmin, box[c][3..5] = child c bmax   (exact doubles)
//     ref[c]       = child c: PairNode index (branch) or first TriRecord (leaf)
//     cnt[c]       = kBranch for a branch child, else the leaf's triangle count
//     axis         = the reference node's split axis (near/far order, bvh_accel.cc:818-823)
//   -> one inner-node visit = ONE 128-byte line = eight 16-byte vector loads,
//      and yields both child tests; half the dependent round trips of the
//      reference layout for the same bytes per box (64 B/box incl. metadata).
//
//   TriRecord (leaf order = indices_ order, so "last visited wins" ties are kept)
//     f32 variant, 48 B : p0.xyz, faceID | p1.xyz, materialID | p2.xyz, pad
//         used when every vertex coordinate is exactly float-representable
//         (always true for OBJ input with scene_scale 1: positions are parsed as
//         float, tiny_obj_loader.cc:696-697); values are widened to double
//         before any arithmetic, so results are unchanged.
//     f64 variant, 80 B : p0, e1 = p1-p0, e2 = p2-p0 as doubles | faceID, materialID
//         the edges are rounded on the host exactly as TriangleIsect rounds them
//         (one IEEE double subtraction each, bvh_accel.cc:600-603), which removes six
//         subtractions (and, against the f32 variant, nine conversions) per test.
//   -> one triangle test = 3 (or 5) 16-byte vector loads, no indirection.
//
// The original faces/vertices/normals/uvs arrays are also resident (verbatim)
// for BuildIntersection (bvh_accel.cc:699-769), which runs once per hit ray.
#ifndef MALLIE_B200_DEVICE_LAYOUT_H_
#define MALLIE_B200_DEVICE_LAYOUT_H_

#include <stdint.h>

namespace mb200 {

static const uint32_t kBranch = 0xFFFFFFFFu;
static const uint32_t kTopBit = 0x80000000u; // ref of a branch child that lives in the staged top-of-tree table

struct alignas(128) PairNode {
  double box[2][6];
  uint32_t ref[2];
  uint32_t cnt[2];
  uint32_t axis;
  uint32_t pad_[3];
};
static_assert(sizeof(PairNode) == 128, "PairNode must be one 128-byte line");

// Development variant (MB200_NODE_LAYOUT=64, kVarNode64): the same pair node in 64 bytes.  Every box coordinate the
// reference builder produces is (vertex coordinate -/+ kEPS) rounded once to double (bvh_accel.cc:285-315); when the
// vertices are float-exact the coordinate is therefore determined by that float, and the kernel rebuilds the exact
// double with one conversion and one addition.  Two 256-bit loads (two L1 wavefronts per lane) instead of four.
struct alignas(64) PairNode64 {
  float box[2][6];  // the float f with (double)f -/+ kEPS == the exact box coordinate (checked when the copy is made)
  uint32_t ref[2];
  uint16_t cnt[2];  // 0xFFFF = branch
  uint32_t axis;
};
static_assert(sizeof(PairNode64) == 64, "PairNode64");

struct alignas(16) TriRecordF32 {
  float p0[3];
  uint32_t face;
  float p1[3];
  uint32_t mat;
  float p2[3];
  uint32_t pad_;
};
static_assert(sizeof(TriRecordF32) == 48, "TriRecordF32");

struct alignas(16) TriRecordF64 {
  double p0[3];
  double e1[3]; // p1 - p0, rounded as TriangleIsect rounds it (bvh_accel.cc:600-601)
  double e2[3]; // p2 - p0
  uint32_t face;
  uint32_t mat;
};
static_assert(sizeof(TriRecordF64) == 80, "TriRecordF64");

// Traversal copies of the records, padded to a whole number of 32-byte sectors so that one triangle is two or
// three 256-bit loads (the traversal kernel is bound by L1 wavefronts and issue slots, not by bytes: DESIGN.md §5):
//   kTriF32x64  64 B : the f32 record + 16 B of padding                      -> 2 x LDG.256 instead of 3 x LDG.128
//   kTriF64x96  96 B : p0, e1, e2 as doubles | faceID, materialID | padding  -> 3 x LDG.256, no conversions and no
//                      edge subtractions in the kernel (e = (double)p1 - (double)p0 is formed once, by the layout
//                      kernel, with the same single IEEE subtraction TriangleIsect performs, bvh_accel.cc:600-603)
//   kTriWoop    96 B : Woop's affine map into the unit triangle (rows r1, r2, r3 = n and offsets b = -r . p0, doubles);
//                      faceID / materialID are read from the canonical record when a hit is accepted.  NOT exact: its
//                      roundings differ from TriangleIsect's, so hit records are not bit-identical to the reference.
//                      Development builds only, to measure what north_star's "Woop-packed triangles" would buy and
//                      cost (tools/woop_report.py).
enum TriKind : int { kTriF32 = 0, kTriF64 = 1, kTriF32x64 = 2, kTriF64x96 = 3, kTriWoop = 4 };
#ifdef __CUDACC__
__host__ __device__
#endif
    inline unsigned
    tri_kind_bytes(int kind) {
  return kind == kTriF32 ? 48u : kind == kTriF64 ? 80u : kind == kTriF32x64 ? 64u : 96u;
}

// ---- wavefront buffers of the frame kernels (kernels.cu) ------------------------------------------------
// A queued secondary ray (shadow ray or path continuation): 64 B = four 16-byte vector loads.
//   item  = the work item (sample of a pixel) the ray belongs to
//   value = shadow rays: the radiance the sample receives if the ray is NOT occluded
struct alignas(16) QRay {
  double org[3];
  double dir[3];
  double tmax;
  uint32_t item;
  float value;
};
static_assert(sizeof(QRay) == 64, "QRay");

// Per-sample state of PathTrace (render.cc:381-456) between wavefronts.  Radiance and throughput are
// scalars: every Material the reference can produce is the grey default (scene.h:58-65), so r = g = b.
struct alignas(16) PathState {
  uint32_t rng[4];  // xorshift128 state (render.cc:137-168)
  double throughput;
  double radiance;
  uint32_t cur_mat; // Intersection::materialID as PathTrace's isect variable holds it (stale after a miss)
  uint32_t pad_[3];
};
static_assert(sizeof(PathState) == 48, "PathState");

// Everything a kernel needs, passed by value as a __grid_constant__ parameter.
struct SceneView {
  const PairNode *nodes;     // [num_pair_nodes]
  const void *tris;          // TriRecordF32[] or TriRecordF64[]  (num_tris)
  double root_box[6];        // reference node 0 bounds
  uint32_t root_ref;         // as PairNode::ref
  uint32_t root_cnt;         // as PairNode::cnt
  uint32_t num_pair_nodes;
  uint32_t num_tris;
  int empty;                 // no nodes at all: every ray misses
  int tri_f32;               // 1 = TriRecordF32
  const void *trav_tris;     // what the traversal kernels read: `tris`, or a padded copy of it (TriKind)
  int tri_kind;              // TriKind of trav_tris
  // Development variant (MB200_TOP_NODES=K, kVarTopSmem): the first K pair nodes in breadth-first order, child refs
  // that stay inside the table re-numbered and tagged with kTopBit; every CTA stages the table into shared memory
  // with cp.async.bulk at kernel start (trace_sm.cuh).
  const PairNode *top_nodes;
  uint32_t top_count;
  const PairNode64 *nodes64; // development variant kVarNode64 (null: not available for this scene)
  // Octant copies of the pair nodes (kVarOctNodes; null: none): copy s (bit a of s = direction component a is negative)
  // at nodes_oct + s * num_pair_nodes.  In copy s every box is stored as (near x, near y, near z, far x, far y, far z)
  // for rays of that octant -- what IntersectRayAABB selects with dirSign (bvh_accel.cc:556-561) -- and the children
  // are ordered (near, far) as Traverse orders them with dirSign[axis] (bvh_accel.cc:818-823), so the inner step
  // needs no per-axis selects and no axis: the same values reach the same arithmetic.  Branch refs inside a copy are
  // absolute (s * num_pair_nodes + index), so a ray enters its copy at the root and never leaves it.
  const PairNode *nodes_oct;
  // verbatim mesh (mesh.h:7-18) for BuildIntersection
  const double *vertices;    // [3*nv]
  const uint32_t *faces;     // [3*nf]
  const double *fv_normals;  // [9*nf] or null
  const double *fv_uvs;      // [6*nf] or null
  uint32_t num_vertices;
  uint32_t num_faces;
};

} // namespace mb200

#endif

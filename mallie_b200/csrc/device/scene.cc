// scene.cc -- builds the device-resident scene from Mallie's host data:
// validates the reference-layout BVH (so a malformed tree can never hang a GPU),
// converts it to the children-in-parent PairNode layout + leaf-ordered triangle
// records (layout.h) and uploads it together with the verbatim mesh arrays.
#include "scene.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mb200 {

namespace {

bool all_float_exact(const double *v, size_t n) {
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok) if (n > (1u << 16))
  for (long i = 0; i < (long)n; i++) {
    const double x = v[i];
    ok = ok && std::isfinite(x) && (double)(float)x == x;
  }
  return ok;
}

} // namespace

bool choose_tri_f32(const double *vertices, size_t count) {
  if (const char *fmt = getenv("MB200_TRI_FORMAT")) { // development knob: "f64" forces the 80-byte edge records
    if (!strcmp(fmt, "f64")) return false;
  }
  return all_float_exact(vertices, count);
}

int relayout_bvh(Relayout &out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
                 const uint32_t *material_ids, const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices,
                 size_t nindices, std::string *err) {
  auto fail = [&](const char *m) {
    if (err) *err = m;
    return (int)MB200_ERR_INVALID_ARG;
  };
  out = Relayout();
  if (nnodes == 0) return MB200_OK; // empty scene: every ray misses
  const char *tenv = getenv("MB200_BUILD_TIMING");
  const bool timing = tenv && atoi(tenv) != 0;
  auto t_prev = std::chrono::steady_clock::now();
  auto stage = [&](const char *name) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mb200 relayout] %-10s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  if (!nodes || !indices || !vertices || !faces) return fail("null mesh/BVH array");
  if (nnodes >= 0xFFFFFFF0ull || nindices >= 0xFFFFFFF0ull) return fail("BVH too large for 32-bit references");
  {
    bool idx_ok = true, face_ok = true;
#pragma omp parallel for schedule(static) reduction(&& : idx_ok) if (nindices > (1u << 16))
    for (long i = 0; i < (long)nindices; i++) idx_ok = idx_ok && indices[i] < nfaces;
#pragma omp parallel for schedule(static) reduction(&& : face_ok) if (nfaces > (1u << 15))
    for (long i = 0; i < (long)(3 * nfaces); i++) face_ok = face_ok && faces[i] < nverts;
    if (!idx_ok) return fail("BVH index array references a face out of range");
    if (!face_ok) return fail("face references a vertex out of range");
  }

  stage("validate");
  // ---- pass 1: walk the tree (pre-order, explicit stack), validate, number the branches
  std::vector<uint32_t> pair_of(nnodes, 0xFFFFFFFFu);
  std::vector<unsigned char> seen(nnodes, 0);
  struct Item {
    uint32_t node;
    int depth;
  };
  std::vector<Item> stack;
  stack.push_back({0u, 0});
  uint32_t npairs = 0;
  int max_depth = 0;
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (it.node >= nnodes) return fail("BVH child index out of range");
    if (seen[it.node]) return fail("BVH is not a tree (node reachable twice)");
    seen[it.node] = 1;
    if (it.depth > max_depth) max_depth = it.depth;
    const mb200_bvh_node &nd = nodes[it.node];
    if (nd.flag == 0) {
      if (nd.axis < 0 || nd.axis > 2) return fail("BVH branch node has an invalid split axis");
      pair_of[it.node] = npairs++;
      stack.push_back({nd.data[1], it.depth + 1});
      stack.push_back({nd.data[0], it.depth + 1});
    } else {
      if ((size_t)nd.data[1] + (size_t)nd.data[0] > nindices) return fail("BVH leaf range exceeds the index array");
    }
  }
  if (max_depth > 500) return fail("BVH deeper than 500 levels (reference stack is 512, bvh_accel.cc:548)");

  stage("walk");
  // ---- pass 2: emit PairNodes
  out.empty = false;
  out.depth = max_depth;
  if (!out.pairs.resize(npairs)) return fail("out of host memory");
  auto child_ref = [&](uint32_t node, uint32_t &ref, uint32_t &cnt) {
    const mb200_bvh_node &c = nodes[node];
    if (c.flag == 0) {
      ref = pair_of[node];
      cnt = kBranch;
    } else {
      ref = c.data[1];
      cnt = c.data[0];
    }
  };
  child_ref(0, out.root_ref, out.root_cnt);
#pragma omp parallel for schedule(static) if (nnodes > (1u << 14))
  for (long i = 0; i < (long)nnodes; i++) {
    if (!seen[i] || nodes[i].flag != 0) continue;
    PairNode &p = out.pairs[pair_of[i]];
    memset(&p, 0, sizeof(p));
    for (int c = 0; c < 2; c++) {
      const mb200_bvh_node &ch = nodes[nodes[i].data[c]];
      for (int k = 0; k < 3; k++) {
        p.box[c][k] = ch.bmin[k];
        p.box[c][3 + k] = ch.bmax[k];
      }
      child_ref(nodes[i].data[c], p.ref[c], p.cnt[c]);
    }
    p.axis = (uint32_t)nodes[i].axis;
  }

  stage("pairs");
  // ---- triangle records in indices_ order
  out.f32 = choose_tri_f32(vertices, 3 * nverts);
  if (out.f32) {
    if (!out.tris32.resize(nindices)) return fail("out of host memory");
#pragma omp parallel for schedule(static) if (nindices > (1u << 14))
    for (long i = 0; i < (long)nindices; i++) {
      const uint32_t f = indices[i];
      TriRecordF32 &t = out.tris32[i];
      const double *a = vertices + 3 * (size_t)faces[3 * (size_t)f + 0];
      const double *b = vertices + 3 * (size_t)faces[3 * (size_t)f + 1];
      const double *c = vertices + 3 * (size_t)faces[3 * (size_t)f + 2];
      for (int k = 0; k < 3; k++) t.p0[k] = (float)a[k], t.p1[k] = (float)b[k], t.p2[k] = (float)c[k];
      t.face = f;
      t.mat = material_ids ? material_ids[f] : 0xFFFFFFFFu; // bvh_accel.cc:685-689
      t.pad_ = 0;
    }
  } else {
    if (!out.tris64.resize(nindices)) return fail("out of host memory");
#pragma omp parallel for schedule(static) if (nindices > (1u << 14))
    for (long i = 0; i < (long)nindices; i++) {
      const uint32_t f = indices[i];
      TriRecordF64 &t = out.tris64[i];
      const double *a = vertices + 3 * (size_t)faces[3 * (size_t)f + 0];
      const double *b = vertices + 3 * (size_t)faces[3 * (size_t)f + 1];
      const double *c = vertices + 3 * (size_t)faces[3 * (size_t)f + 2];
      for (int k = 0; k < 3; k++) t.p0[k] = a[k], t.e1[k] = b[k] - a[k], t.e2[k] = c[k] - a[k];
      t.face = f;
      t.mat = material_ids ? material_ids[f] : 0xFFFFFFFFu;
    }
  }
  stage("triangles");
  return MB200_OK;
}

namespace {

struct Uploader {
  mb200_scene *s;
  std::string *err;
  int status = MB200_OK;
  template <typename T> const T *put(const T *host, size_t count) {
    if (status != MB200_OK || count == 0 || !host) return nullptr;
    void *d = nullptr;
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) {
      status = (e == cudaErrorMemoryAllocation) ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA;
      if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
      return nullptr;
    }
    s->allocs.push_back(d);
    s->device_bytes += bytes;
    e = cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, s->stream);
    if (e != cudaSuccess) {
      status = MB200_ERR_CUDA;
      if (err) *err = std::string("cudaMemcpy H2D: ") + cudaGetErrorString(e);
      return nullptr;
    }
    return reinterpret_cast<const T *>(d);
  }
};

} // namespace

// Device checks, the handle and its stream: the common first step of scene_create / scene_build_device.
int scene_open(mb200_scene **out, int device, std::string *err) {
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    if (err) *err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return MB200_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    if (err) *err = "device ordinal out of range";
    return MB200_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    if (err) *err = "device is not sm_100 class (this library ships sm_100a code only)";
    return MB200_ERR_NO_DEVICE;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    if (err) *err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return MB200_ERR_CUDA;
  }
  mb200_scene *s = new mb200_scene;
  s->device = device;
  memset(&s->view, 0, sizeof(s->view));
  s->view.empty = 1;
  if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    if (err) *err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete s;
    return MB200_ERR_CUDA;
  }
  *out = s;
  return MB200_OK;
}

// Which triangle records the traversal kernels read.  Default: the canonical ones (48-byte float-exact vertices
// or 80-byte p0 + edges).  MB200_TRI_LAYOUT=64 | 96 (development builds: the kernels for them are compiled with
// make DEV=1) makes a padded traversal copy on the device -- 64 B: the f32 record as two 256-bit loads; 96 B:
// p0 + edges in double as three 256-bit loads.
static int wanted_tri_kind(const SceneView &v) {
  const char *e = getenv("MB200_TRI_LAYOUT");
  const int want = e ? atoi(e) : 0;
  if (want == 96) return kTriF64x96;
  if (e && !strcmp(e, "woop")) return kTriWoop; // not bit-exact: measurement only (layout.h)
  if (want == 64 && v.tri_f32) return kTriF32x64;
  return v.tri_f32 ? kTriF32 : kTriF64;
}

// The work counter / statistics words and the traversal copy of the triangle records, then wait for the uploads.
int scene_finish(mb200_scene *s, std::string *err) {
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, 16 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return MB200_ERR_CUDA;
  }
  s->allocs.push_back(d);
  s->d_work = (unsigned long long *)d;
  s->d_counters = s->d_work + 1; // [8] traversal counters / render stats
  cudaMemsetAsync(d, 0, 16 * sizeof(unsigned long long), s->stream);
  SceneView &v = s->view;
  v.trav_tris = v.tris;
  v.tri_kind = v.tri_f32 ? kTriF32 : kTriF64;
  const int kind = wanted_tri_kind(v);
  if (!v.empty && v.num_tris && kind != v.tri_kind) {
    void *pad = nullptr;
    const size_t bytes = (size_t)v.num_tris * tri_kind_bytes(kind);
    if ((e = cudaMalloc(&pad, bytes)) == cudaSuccess) {
      s->allocs.push_back(pad);
      s->device_bytes += bytes;
      e = launch_pad_tris(v.tris, v.tri_f32, v.num_tris, kind, pad, s->stream);
    }
    if (e != cudaSuccess) {
      if (err) *err = std::string("traversal triangle records: ") + cudaGetErrorString(e);
      return e == cudaErrorMemoryAllocation ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA;
    }
    v.trav_tris = pad, v.tri_kind = kind;
  }
  // octant copies of the pair nodes (layout.h: nodes_oct)
  v.nodes_oct = nullptr;
  // (production layout; MB200_NODE_OCT=0 keeps the canonical nodes only.  8 x 128 B per branch node: 99 MB for the
  // 1 M-triangle scene, 1 GB for 10 M triangles; without the memory for it the scene walks the canonical nodes.)
  const char *no = getenv("MB200_NODE_OCT");
  if (!(no && atoi(no) == 0) && !v.empty && v.num_pair_nodes > 0 && (size_t)v.num_pair_nodes * 8 < 0xFFFFFFFFull) {
    void *d8 = nullptr;
    const size_t bytes = (size_t)v.num_pair_nodes * 8 * sizeof(PairNode);
    if ((e = cudaMalloc(&d8, bytes)) == cudaSuccess) {
      s->allocs.push_back(d8);
      s->device_bytes += bytes;
      if ((e = launch_octant_nodes(v.nodes, v.num_pair_nodes, (PairNode *)d8, s->stream)) != cudaSuccess) {
        if (err) *err = std::string("octant node copies: ") + cudaGetErrorString(e);
        return MB200_ERR_CUDA;
      }
      v.nodes_oct = (const PairNode *)d8;
    } else {
      cudaGetLastError(); // out of memory: not an error, the canonical nodes do
    }
  }
  // 64-byte pair nodes (development builds; layout.h: PairNode64)
  v.nodes64 = nullptr;
  const char *nl = getenv("MB200_NODE_LAYOUT");
  if (nl && atoi(nl) == 64 && !v.empty && v.num_pair_nodes > 0 && v.tri_f32) {
    void *d64 = nullptr;
    int *d_bad = nullptr, bad = 1;
    if ((e = cudaMalloc(&d64, (size_t)v.num_pair_nodes * sizeof(PairNode64))) == cudaSuccess &&
        (e = cudaMalloc((void **)&d_bad, sizeof(int))) == cudaSuccess &&
        (e = cudaMemsetAsync(d_bad, 0, sizeof(int), s->stream)) == cudaSuccess &&
        (e = launch_pack_nodes64(v.nodes, v.num_pair_nodes, (PairNode64 *)d64, d_bad, s->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream)) == cudaSuccess)
      e = cudaStreamSynchronize(s->stream);
    if (d_bad) cudaFree(d_bad);
    if (e != cudaSuccess) {
      if (err) *err = std::string("64-byte pair nodes: ") + cudaGetErrorString(e);
      return MB200_ERR_CUDA;
    }
    if (bad) cudaFree(d64); // some coordinate is not float -/+ kEPS: keep the 128-byte nodes
    else s->allocs.push_back(d64), s->device_bytes += (size_t)v.num_pair_nodes * sizeof(PairNode64), v.nodes64 = (const PairNode64 *)d64;
  }
  // top-of-tree table for the shared-memory staging variant (development builds; layout.h)
  v.top_nodes = nullptr, v.top_count = 0;
  const char *tn = getenv("MB200_TOP_NODES");
  const int want_top = tn ? atoi(tn) : 0;
  if (want_top > 0 && !v.empty && v.root_cnt == kBranch && v.num_pair_nodes > 0) {
    std::vector<PairNode> all(v.num_pair_nodes);
    if ((e = cudaMemcpyAsync(all.data(), v.nodes, all.size() * sizeof(PairNode), cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s->stream)) != cudaSuccess) {
      if (err) *err = std::string("top-of-tree table: ") + cudaGetErrorString(e);
      return MB200_ERR_CUDA;
    }
    std::vector<uint32_t> order; // breadth-first from the root pair
    std::vector<uint32_t> slot(v.num_pair_nodes, 0xFFFFFFFFu);
    order.push_back(v.root_ref);
    slot[v.root_ref] = 0;
    for (size_t head = 0; head < order.size() && order.size() < (size_t)want_top; head++)
      for (int c = 0; c < 2 && order.size() < (size_t)want_top; c++) {
        const PairNode &n = all[order[head]];
        if (n.cnt[c] == kBranch && slot[n.ref[c]] == 0xFFFFFFFFu) {
          slot[n.ref[c]] = (uint32_t)order.size();
          order.push_back(n.ref[c]);
        }
      }
    std::vector<PairNode> top(order.size());
    for (size_t i = 0; i < order.size(); i++) {
      top[i] = all[order[i]];
      for (int c = 0; c < 2; c++)
        if (top[i].cnt[c] == kBranch && slot[top[i].ref[c]] != 0xFFFFFFFFu) top[i].ref[c] = slot[top[i].ref[c]] | kTopBit;
    }
    void *d_top = nullptr;
    if ((e = cudaMalloc(&d_top, top.size() * sizeof(PairNode))) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_top, top.data(), top.size() * sizeof(PairNode), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s->stream)) != cudaSuccess) {
      if (err) *err = std::string("top-of-tree table: ") + cudaGetErrorString(e);
      return MB200_ERR_CUDA;
    }
    s->allocs.push_back(d_top);
    v.top_nodes = (const PairNode *)d_top, v.top_count = (uint32_t)top.size();
  }
  if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) {
    if (err) *err = std::string("upload: ") + cudaGetErrorString(e);
    return MB200_ERR_CUDA;
  }
  return MB200_OK;
}

int scene_create(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                 size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                 const mb200_bvh_node *nodes, size_t nnodes, const uint32_t *indices, size_t nindices,
                 std::string *err) {
  mb200_scene *s = nullptr;
  int st = scene_open(&s, device, err);
  if (st != MB200_OK) return st;
  *out = nullptr;
  Relayout rl;
  st = relayout_bvh(rl, vertices, nverts, faces, nfaces, material_ids, nodes, nnodes, indices, nindices, err);
  if (st != MB200_OK) {
    scene_destroy(s);
    return st;
  }
  Uploader up{s, err};
  SceneView &v = s->view;
  memset(&v, 0, sizeof(v));
  v.empty = rl.empty ? 1 : 0;
  v.tri_f32 = rl.f32 ? 1 : 0;
  v.num_vertices = (uint32_t)nverts;
  v.num_faces = (uint32_t)nfaces;
  if (!rl.empty) {
    v.nodes = up.put(rl.pairs.data(), rl.pairs.size());
    v.tris = rl.f32 ? (const void *)up.put(rl.tris32.data(), rl.tris32.size())
                    : (const void *)up.put(rl.tris64.data(), rl.tris64.size());
    v.num_pair_nodes = (uint32_t)rl.pairs.size();
    v.num_tris = (uint32_t)nindices;
    v.root_ref = rl.root_ref;
    v.root_cnt = rl.root_cnt;
    for (int k = 0; k < 3; k++) {
      v.root_box[k] = nodes[0].bmin[k];
      v.root_box[3 + k] = nodes[0].bmax[k];
      s->root_bmin[k] = nodes[0].bmin[k];
      s->root_bmax[k] = nodes[0].bmax[k];
    }
    v.vertices = up.put(vertices, 3 * nverts);
    v.faces = up.put(faces, 3 * nfaces);
    v.fv_normals = fv_normals ? up.put(fv_normals, 9 * nfaces) : nullptr;
    v.fv_uvs = fv_uvs ? up.put(fv_uvs, 6 * nfaces) : nullptr;
  }
  s->tree_depth = rl.depth;
  s->stack_cap = rl.depth + 2;

  if (up.status == MB200_OK) up.status = scene_finish(s, err);
  if (up.status != MB200_OK) {
    scene_destroy(s);
    return up.status;
  }
  *out = s;
  return MB200_OK;
}

// A replica of `src` on GPU `device`: the resident arrays are copied device to device (NVLink when the GPUs are
// peers), nothing is rebuilt or re-laid out on the host.
int scene_clone(mb200_scene **out, mb200_scene *src, int device, std::string *err) {
  *out = nullptr;
  cudaError_t e = cudaSetDevice(src->device);
  if (e == cudaSuccess) e = cudaStreamSynchronize(src->stream);
  if (e != cudaSuccess) {
    if (err) *err = std::string("source scene: ") + cudaGetErrorString(e);
    return MB200_ERR_CUDA;
  }
  mb200_scene *s = nullptr;
  int st = scene_open(&s, device, err);
  if (st != MB200_OK) return st;
  if (device != src->device) { // direct NVLink copies need the peer mapping; without it the copy is staged through the host
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
      const cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0); // current device is `device` (scene_open)
      (void)pe;                                                           // already enabled is fine
    }
    cudaGetLastError();
  }
  const SceneView &sv = src->view;
  SceneView &v = s->view;
  v = sv;
  v.nodes = nullptr, v.tris = nullptr, v.trav_tris = nullptr, v.vertices = nullptr, v.faces = nullptr, v.fv_normals = nullptr, v.fv_uvs = nullptr;
  auto copy = [&](const void *from, size_t bytes) -> const void * {
    if (st != MB200_OK || !from || bytes == 0) return nullptr;
    void *d = nullptr;
    if ((e = cudaMalloc(&d, bytes)) != cudaSuccess) {
      st = (e == cudaErrorMemoryAllocation) ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA;
      if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
      return nullptr;
    }
    s->allocs.push_back(d);
    s->device_bytes += bytes;
    if ((e = cudaMemcpyPeerAsync(d, device, from, src->device, bytes, s->stream)) != cudaSuccess) {
      st = MB200_ERR_CUDA;
      if (err) *err = std::string("cudaMemcpyPeer: ") + cudaGetErrorString(e);
      return nullptr;
    }
    return d;
  };
  if (!sv.empty) {
    const size_t nf = sv.num_faces, nv = sv.num_vertices;
    v.nodes = (const PairNode *)copy(sv.nodes, (size_t)sv.num_pair_nodes * sizeof(PairNode));
    v.tris = copy(sv.tris, (size_t)sv.num_tris * (sv.tri_f32 ? sizeof(TriRecordF32) : sizeof(TriRecordF64)));
    v.vertices = (const double *)copy(sv.vertices, 3 * nv * sizeof(double));
    v.faces = (const uint32_t *)copy(sv.faces, 3 * nf * sizeof(uint32_t));
    v.fv_normals = (const double *)copy(sv.fv_normals, 9 * nf * sizeof(double));
    v.fv_uvs = (const double *)copy(sv.fv_uvs, 6 * nf * sizeof(double));
  }
  for (int k = 0; k < 3; k++) s->root_bmin[k] = src->root_bmin[k], s->root_bmax[k] = src->root_bmax[k];
  s->tree_depth = src->tree_depth;
  s->stack_cap = src->stack_cap;
  if (st == MB200_OK) st = scene_finish(s, err);
  if (st != MB200_OK) {
    scene_destroy(s);
    return st;
  }
  *out = s;
  return MB200_OK;
}

void scene_destroy(mb200_scene *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->pipe.aux) cudaStreamSynchronize(s->pipe.aux);
  if (s->pipe.copy) cudaStreamSynchronize(s->pipe.copy);
  for (void *p : s->allocs) cudaFree(p);
  mb200_scene::Staging *sts[4] = {&s->in0, &s->in1, &s->out0, &s->out1};
  for (auto *st : sts) {
    if (st->pinned) cudaFreeHost(st->pinned);
    if (st->dev) cudaFree(st->dev);
  }
  s->timer.release();
  s->pipe.release();
  frame_scratch_release(s->frame_scratch);
  frame_scratch_release(s->hit_scratch);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

} // namespace mb200

// gather.cu -- the one collective of the path: NCCL all-gather of the framebuffer for frames that are split by
// row bands over one process per GPU (the reference has no equivalent, its MPI is an
// init/finalize stub: main.cc:213-236).
//
// Every rank renders its bands into a compact buffer that IS the NCCL send buffer (mb200_render_params
// band_compact), one ncclAllGather on the scene's stream moves the bands, and k_deinterleave_rows -- this
// library's kernel, on the same stream -- writes the rows into place.  NCCL is bound at run time
// (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded, so the two never
// disagree about versions); a host without NCCL gets MB200_ERR_UNSUPPORTED from mb200_comm_*, nothing else changes.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "kernels.h"
#include "scene.h"

namespace mb200 {
int capi_set_error(int code, const std::string &msg); // capi.cc: records the calling thread's error string
}

namespace {

struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclCommCount) CommCount = nullptr;
  decltype(&ncclCommUserRank) CommUserRank = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string why;
};

NcclApi *nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // MB200_NCCL_LIB names the library explicitly.  Otherwise the soname: a copy the process has already loaded
    // (PyTorch's bundled one after `import torch`) is reused by the dynamic loader.  A process that loads PyTorch
    // AFTER its first mb200_comm_* call would hand torch this (possibly older) copy: import torch first there.
    const char *forced = getenv("MB200_NCCL_LIB");
    for (const char *name : {forced, "libnccl.so.2", "libnccl.so"}) {
      if (!name || !*name) continue;
      api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.why = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found");
      return;
    }
#define MB200_SYM(field, sym)                                         \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym)); \
  if (!api.field) api.why = std::string("NCCL lacks ") + sym;
    MB200_SYM(GetUniqueId, "ncclGetUniqueId")
    MB200_SYM(CommInitRank, "ncclCommInitRank")
    MB200_SYM(CommDestroy, "ncclCommDestroy")
    MB200_SYM(AllGather, "ncclAllGather")
    MB200_SYM(GetErrorString, "ncclGetErrorString")
    MB200_SYM(CommCount, "ncclCommCount")
    MB200_SYM(CommUserRank, "ncclCommUserRank")
    MB200_SYM(GetVersion, "ncclGetVersion")
#undef MB200_SYM
  });
  return api.why.empty() ? &api : nullptr;
}

std::string nccl_why() { return "NCCL is not available in this process (libnccl.so.2 could not be loaded or lacks a symbol)"; }

// recv = [ranks][pad_rows][row_floats]: rank r's compact band buffer (band b of the image belongs to rank b % ranks,
// its rows follow each other in the compact buffer).  One thread per float4 of the output image.
__global__ void __launch_bounds__(256) k_deinterleave_rows(const float4 *__restrict__ recv, float4 *__restrict__ image,
                                                           int height, int row_vec4, int band_rows, int ranks,
                                                           int pad_rows) {
  const size_t total = (size_t)height * row_vec4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / row_vec4), x = (int)(i - (size_t)y * row_vec4);
    const int band = y / band_rows, r = band % ranks;
    const int local = (band / ranks) * band_rows + (y - band * band_rows);
    image[i] = __ldg(recv + ((size_t)r * pad_rows + local) * row_vec4 + x);
  }
}

__global__ void __launch_bounds__(256) k_deinterleave_rows_scalar(const float *__restrict__ recv, float *__restrict__ image,
                                                                  int height, int row_floats, int band_rows, int ranks,
                                                                  int pad_rows) {
  const size_t total = (size_t)height * row_floats;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / row_floats), x = (int)(i - (size_t)y * row_floats);
    const int band = y / band_rows, r = band % ranks;
    const int local = (band / ranks) * band_rows + (y - band * band_rows);
    image[i] = __ldg(recv + ((size_t)r * pad_rows + local) * row_floats + x);
  }
}

__global__ void __launch_bounds__(256) k_fill_int(int *__restrict__ dst, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

bool device_pointer(const void *p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool pinned_pointer(const void *p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

} // namespace

struct mb200_comm {
  mb200_scene *scene = nullptr;
  ncclComm_t comm = nullptr;
  bool owned = false;
  int nranks = 1, rank = 0;
  // grown on demand: this rank's padded band buffer (the send buffer), the gathered bands, the assembled frame
  float *send = nullptr, *recv = nullptr, *full = nullptr;
  int *cnt = nullptr;
  size_t send_cap = 0, recv_cap = 0, full_cap = 0, cnt_cap = 0;
  void *pinned = nullptr; // staging for pageable host destinations
  size_t pinned_cap = 0;
};

namespace {

int grow(void **p, size_t *cap, size_t bytes, cudaStream_t s) {
  if (*cap >= bytes) return MB200_OK;
  if (*p) {
    if (cudaStreamSynchronize(s) != cudaSuccess) return MB200_ERR_CUDA;
    cudaFree(*p);
    *p = nullptr, *cap = 0;
  }
  const cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    *p = nullptr;
    return mb200::capi_set_error(e == cudaErrorMemoryAllocation ? MB200_ERR_OUT_OF_MEMORY : MB200_ERR_CUDA,
                                 std::string("gather buffers: ") + cudaGetErrorString(e));
  }
  *cap = bytes;
  return MB200_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  NcclApi *n = nccl();
  return mb200::capi_set_error(MB200_ERR_CUDA, std::string(what) + ": " + (n ? n->GetErrorString(r) : "NCCL error"));
}

int max_band_rows(int height, int band_rows, int ranks) {
  int best = 0;
  for (int r = 0; r < ranks; r++) {
    const int rows = mb200::band_rows_owned(height, band_rows, ranks, r);
    if (rows > best) best = rows;
  }
  return best;
}

} // namespace

extern "C" {

int mb200_comm_unique_id(unsigned char id[MB200_COMM_ID_BYTES]) {
  if (!id) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "id is null");
  static_assert(MB200_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  NcclApi *n = nccl();
  if (!n) return mb200::capi_set_error(MB200_ERR_UNSUPPORTED, nccl_why());
  ncclUniqueId u;
  const ncclResult_t r = n->GetUniqueId(&u);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
  return MB200_OK;
}

static int comm_new(mb200_comm **out, mb200_scene *scene, ncclComm_t c, bool owned, int nranks, int rank) {
  mb200_comm *m = new mb200_comm;
  m->scene = scene, m->comm = c, m->owned = owned, m->nranks = nranks, m->rank = rank;
  *out = m;
  return MB200_OK;
}

int mb200_comm_init(mb200_comm **out, mb200_scene *scene, int nranks, int rank, const unsigned char id[MB200_COMM_ID_BYTES]) {
  if (!out) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if (!scene || !id || nranks < 1 || rank < 0 || rank >= nranks)
    return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "bad communicator arguments");
  NcclApi *n = nccl();
  if (!n) return mb200::capi_set_error(MB200_ERR_UNSUPPORTED, nccl_why());
  if (cudaSetDevice(scene->device) != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "cudaSetDevice failed");
  ncclUniqueId u;
  memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  const ncclResult_t r = n->CommInitRank(&c, nranks, u, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  return comm_new(out, scene, c, true, nranks, rank);
}

int mb200_comm_adopt(mb200_comm **out, mb200_scene *scene, void *nccl_comm) {
  if (!out) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if (!scene || !nccl_comm) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "null argument");
  NcclApi *n = nccl();
  if (!n) return mb200::capi_set_error(MB200_ERR_UNSUPPORTED, nccl_why());
  int count = 0, rank = 0;
  ncclResult_t r = n->CommCount((ncclComm_t)nccl_comm, &count);
  if (r == ncclSuccess) r = n->CommUserRank((ncclComm_t)nccl_comm, &rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommCount / ncclCommUserRank");
  return comm_new(out, scene, (ncclComm_t)nccl_comm, false, count, rank);
}

int mb200_comm_size(const mb200_comm *c) { return c ? c->nranks : 0; }
int mb200_comm_rank(const mb200_comm *c) { return c ? c->rank : -1; }

void mb200_comm_destroy(mb200_comm *c) {
  if (!c) return;
  if (c->scene) {
    cudaSetDevice(c->scene->device);
    cudaStreamSynchronize(c->scene->stream);
  }
  NcclApi *n = nccl();
  if (c->owned && c->comm && n) n->CommDestroy(c->comm);
  for (void *p : {(void *)c->send, (void *)c->recv, (void *)c->full, (void *)c->cnt})
    if (p) cudaFree(p);
  if (c->pinned) cudaFreeHost(c->pinned);
  delete c;
}

// The band layout of a frame split over the communicator's ranks: rows of this rank, rows of the fullest rank.
static int band_layout(const mb200_comm *c, int height, int band_rows, int *local_rows, int *pad_rows) {
  if (band_rows < 4 || band_rows % 4 != 0) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "band_rows must be a positive multiple of 4");
  *local_rows = mb200::band_rows_owned(height, band_rows, c->nranks, c->rank);
  *pad_rows = max_band_rows(height, band_rows, c->nranks);
  return MB200_OK;
}

int mb200_gather_framebuffer(mb200_comm *c, int width, int height, int channels, int band_rows, const float *d_bands,
                             float *image) {
  if (!c || width <= 0 || height <= 0 || channels <= 0) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "bad argument");
  NcclApi *n = nccl();
  if (!n) return mb200::capi_set_error(MB200_ERR_UNSUPPORTED, nccl_why());
  mb200_scene *s = c->scene;
  if (cudaSetDevice(s->device) != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "cudaSetDevice failed");
  int local_rows = 0, pad_rows = 0, rc;
  if ((rc = band_layout(c, height, band_rows, &local_rows, &pad_rows)) != MB200_OK) return rc;
  const size_t row_floats = (size_t)width * channels;
  const size_t send_floats = (size_t)pad_rows * row_floats, full_floats = (size_t)height * row_floats;
  if (local_rows > 0 && !device_pointer(d_bands)) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "d_bands must be a device pointer");
  if ((rc = grow((void **)&c->recv, &c->recv_cap, send_floats * c->nranks * sizeof(float), s->stream)) != MB200_OK) return rc;
  const float *send = d_bands;
  if (d_bands != c->send) { // a caller's buffer holds local_rows rows: NCCL needs equal counts, pad through our send buffer
    if ((rc = grow((void **)&c->send, &c->send_cap, send_floats * sizeof(float), s->stream)) != MB200_OK) return rc;
    if (local_rows > 0 &&
        cudaMemcpyAsync(c->send, d_bands, (size_t)local_rows * row_floats * sizeof(float), cudaMemcpyDeviceToDevice, s->stream) != cudaSuccess)
      return mb200::capi_set_error(MB200_ERR_CUDA, "band copy failed");
    send = c->send;
  }
  const ncclResult_t r = n->AllGather(send, c->recv, send_floats, ncclFloat, c->comm, s->stream);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
  if (!image) return MB200_OK; // this rank does not need the frame
  const bool to_device = device_pointer(image);
  float *dst = image;
  if (!to_device) {
    if ((rc = grow((void **)&c->full, &c->full_cap, full_floats * sizeof(float), s->stream)) != MB200_OK) return rc;
    dst = c->full;
  }
  const int grid = 148 * 8;
  if (row_floats % 4 == 0 && (reinterpret_cast<size_t>(dst) & 15u) == 0)
    k_deinterleave_rows<<<grid, 256, 0, s->stream>>>((const float4 *)c->recv, (float4 *)dst, height, (int)(row_floats / 4),
                                                     band_rows, c->nranks, pad_rows);
  else
    k_deinterleave_rows_scalar<<<grid, 256, 0, s->stream>>>(c->recv, dst, height, (int)row_floats, band_rows, c->nranks, pad_rows);
  mb200::note_launch();
  if (cudaGetLastError() != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "de-interleave launch failed");
  if (to_device) return MB200_OK; // enqueue-only, like the render calls
  void *host = image;
  const bool pinned = pinned_pointer(image);
  if (!pinned) {
    if (c->pinned_cap < full_floats * sizeof(float)) {
      if (c->pinned) cudaFreeHost(c->pinned);
      c->pinned = nullptr, c->pinned_cap = 0;
      if (cudaMallocHost(&c->pinned, full_floats * sizeof(float)) != cudaSuccess)
        return mb200::capi_set_error(MB200_ERR_OUT_OF_MEMORY, "pinned staging for the gathered frame");
      c->pinned_cap = full_floats * sizeof(float);
    }
    host = c->pinned;
  }
  if (cudaMemcpyAsync(host, dst, full_floats * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
      cudaStreamSynchronize(s->stream) != cudaSuccess)
    return mb200::capi_set_error(MB200_ERR_CUDA, "gathered frame: device -> host copy failed");
  if (!pinned) memcpy(image, c->pinned, full_floats * sizeof(float));
  return MB200_OK;
}

int mb200_render_frame_gathered(mb200_comm *c, const mb200_render_params *p, int num_passes, int band_rows, float *image,
                                int *count, mb200_render_stats *stats) {
  if (!c || !p) return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "null argument");
  if (p->width <= 0 || p->height <= 0 || p->x0 != 0 || p->y0 != 0 || p->x1 != p->width || p->y1 != p->height ||
      p->band_rows != 0 || p->pixel_step > 1 || num_passes < 1)
    return mb200::capi_set_error(MB200_ERR_INVALID_ARG, "gathered frames need whole-image parameters without bands / step");
  mb200_scene *s = c->scene;
  if (cudaSetDevice(s->device) != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "cudaSetDevice failed");
  int local_rows = 0, pad_rows = 0, rc;
  if ((rc = band_layout(c, p->height, band_rows, &local_rows, &pad_rows)) != MB200_OK) return rc;
  const size_t W = (size_t)p->width, H = (size_t)p->height;
  if ((rc = grow((void **)&c->send, &c->send_cap, (size_t)pad_rows * W * 3 * sizeof(float), s->stream)) != MB200_OK) return rc;
  if ((rc = grow((void **)&c->cnt, &c->cnt_cap, (size_t)(pad_rows ? pad_rows : 1) * W * sizeof(int), s->stream)) != MB200_OK) return rc;
  mb200_render_params pg = *p;
  pg.band_rows = band_rows, pg.band_count = c->nranks, pg.band_index = c->rank, pg.band_compact = 1;
  if (local_rows > 0) {
    // renders straight into the NCCL send buffer (device pointers: enqueue-only unless stats are read back)
    if ((rc = mb200_render_frame(s, &pg, num_passes, c->send, c->cnt, stats)) != MB200_OK) return rc;
  } else if (stats) {
    memset(stats, 0, sizeof(*stats));
  }
  if (count) { // every pixel of a fresh frame has num_passes samples: nothing to gather
    if (device_pointer(count)) {
      k_fill_int<<<148 * 4, 256, 0, s->stream>>>(count, W * H, num_passes);
      mb200::note_launch();
      if (cudaGetLastError() != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "count fill failed");
    } else {
      std::fill_n(count, W * H, num_passes); // on the host, while the GPU renders
    }
  }
  if ((rc = mb200_gather_framebuffer(c, p->width, p->height, 3, band_rows, c->send, image)) != MB200_OK) return rc;
  return MB200_OK;
}

} // extern "C"

// gather.cu -- the one collective of the path: NCCL all-gather of the framebuffer for frames that are split by
// row bands over one process per GPU (the reference has no equivalent, its MPI is an
// init/finalize stub: main.cc:213-236).
//
// Every rank renders its bands into a compact buffer (mb200_render_params band_compact).  Two ways to exchange them:
//   * peer memory (default when every rank of the communicator can map every other rank's frame buffer: one node,
//     NVLink / NVSwitch): k_exchange_rows -- ONE kernel of this library -- stores the rank's rows straight into place
//     in every rank's frame buffer (cudaIpc-mapped, found once through the communicator), then its last CTA raises
//     this rank's flag in every peer and waits for theirs: no all-gather, no row placement pass, no NCCL launch on
//     the frame's critical path;
//   * NCCL (MB200_GATHER=nccl, or when the mapping fails on any rank): the compact buffer IS the NCCL send buffer, one
//     ncclAllGather on the scene's stream moves the bands, and k_deinterleave_rows writes the rows into place.  NCCL is bound at run time
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
#include <vector>

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

// ---- exchange over peer memory ---------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
constexpr size_t kPeerHeaderBytes = 1024; // [0, 128): flags[kMaxPeers] (u64), [128, 132): CTA counter; frames follow
struct PeerTable {
  char *base[kMaxPeers]; // rank p's buffer as mapped into this process (own rank: the allocation itself)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// bands: this rank's compact rows [local_rows][row_vec4].  Row `local` of the compact buffer is image row
// ((local / band_rows) * ranks + rank) * band_rows + local % band_rows; it is stored there in EVERY rank's frame
// (peer stores over NVLink; the own copy is a local store).  Then the exchange is closed: every CTA fences its stores
// system-wide and counts itself; the last one writes `epoch` into this rank's flag in every peer's header and waits until
// every peer's flag in the own header has reached `epoch` -- at that point all rows of the frame are in local memory.
__global__ void __launch_bounds__(256)
    k_exchange_rows(const float4 *__restrict__ bands, const __grid_constant__ PeerTable t, size_t frame_off, int local_rows,
                    int row_vec4, int band_rows, int ranks, int rank, unsigned long long epoch) {
  const size_t total = (size_t)local_rows * row_vec4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int local = (int)(i / row_vec4), x = (int)(i - (size_t)local * row_vec4);
    const int y = ((local / band_rows) * ranks + rank) * band_rows + local % band_rows;
    const float4 v = __ldg(bands + i);
    const size_t at = (size_t)y * row_vec4 + x;
    for (int p = 0; p < ranks; p++) reinterpret_cast<float4 *>(t.base[p] + frame_off)[at] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  unsigned int *counter = reinterpret_cast<unsigned int *>(t.base[rank] + 128);
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system(); // the other CTAs' stores (ordered before their count) before the flags
  const int p = (int)threadIdx.x;
  if (p < ranks) {
    st_release_sys(reinterpret_cast<unsigned long long *>(t.base[p]) + rank, epoch);
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(t.base[rank]) + p;
    // a peer that died never raises its flag: give up after 5 minutes with a trap (the host sees a CUDA error, not a hang)
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(mine) < epoch) {
      __nanosleep(64);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 300000000000ull) __trap();
    }
  }
  if (threadIdx.x == 0) *counter = 0u; // for the next frame (stream order)
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
  // exchange over peer memory (k_exchange_rows)
  int peer_state = 0;              // 0 not tried for peer_floats, 1 mapped on every rank, -1 not available
  size_t peer_floats = 0;          // floats of one frame the mapping was made for
  char *peer_base[16] = {};        // [rank] = own allocation: header + two frames (alternating by epoch)
  unsigned long long epoch = 0;
  int last_path = 0;               // 1: the last frame went through k_exchange_rows, 0: through NCCL
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

void peer_release(mb200_comm *c) {
  for (int p = 0; p < c->nranks && p < kMaxPeers; p++) {
    if (!c->peer_base[p]) continue;
    if (p == c->rank) cudaFree(c->peer_base[p]);
    else cudaIpcCloseMemHandle(c->peer_base[p]);
    c->peer_base[p] = nullptr;
  }
  c->peer_state = 0, c->peer_floats = 0;
}

// Maps every rank's frame buffer into every rank (collective over the communicator: every rank reaches it with the
// same arguments, at the same point of its call sequence).  Returns 1 when the peer-memory exchange can be used by all
// ranks, -1 when not (then every rank uses NCCL).
int peer_prepare(mb200_comm *c, size_t frame_floats) {
  static const bool nccl_only = [] {
    const char *v = getenv("MB200_GATHER");
    return v && strcmp(v, "nccl") == 0;
  }();
  if (nccl_only || c->nranks < 2 || c->nranks > kMaxPeers) return -1;
  if (c->peer_state != 0 && c->peer_floats == frame_floats) return c->peer_state;
  NcclApi *n = nccl();
  mb200_scene *s = c->scene;
  if (cudaStreamSynchronize(s->stream) != cudaSuccess) return -1;
  peer_release(c);
  struct Record {
    cudaIpcMemHandle_t handle;
    long long ok;
  };
  static_assert(sizeof(Record) == 72, "record");
  std::vector<Record> rec((size_t)c->nranks);
  Record mine;
  memset(&mine, 0, sizeof(mine));
  const size_t bytes = kPeerHeaderBytes + 2 * frame_floats * sizeof(float);
  char *own = nullptr;
  bool ok = cudaMalloc((void **)&own, bytes) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(own, 0, kPeerHeaderBytes, s->stream) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.handle, own) == cudaSuccess;
  mine.ok = ok ? 1 : 0;
  cudaGetLastError();
  // exchange of the handles, then of the outcome of opening them: two tiny all-gathers
  char *dx = nullptr;
  if (cudaMalloc((void **)&dx, sizeof(Record) * (size_t)c->nranks) != cudaSuccess) { // cannot even take part: fatal for the path
    if (own) cudaFree(own);
    cudaGetLastError();
    c->peer_state = -1, c->peer_floats = frame_floats;
    return mb200::capi_set_error(MB200_ERR_OUT_OF_MEMORY, "peer exchange set-up"), -1;
  }
  auto all_gather = [&](Record *host) -> bool {
    if (cudaMemcpyAsync(dx + sizeof(Record) * (size_t)c->rank, &host[c->rank], sizeof(Record), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return false;
    if (n->AllGather(dx + sizeof(Record) * (size_t)c->rank, dx, sizeof(Record), ncclChar, c->comm, s->stream) != ncclSuccess) return false;
    if (cudaMemcpyAsync(host, dx, sizeof(Record) * (size_t)c->nranks, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return false;
    return cudaStreamSynchronize(s->stream) == cudaSuccess;
  };
  rec[c->rank] = mine;
  bool comm_ok = all_gather(rec.data());
  bool all = comm_ok && ok;
  if (comm_ok) {
    c->peer_base[c->rank] = own;
    for (int p = 0; p < c->nranks; p++) {
      if (p == c->rank) continue;
      void *ptr = nullptr;
      if (!rec[p].ok || cudaIpcOpenMemHandle(&ptr, rec[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        all = false;
        continue;
      }
      c->peer_base[p] = (char *)ptr;
    }
    memset(&mine, 0, sizeof(mine));
    mine.ok = all ? 1 : 0;
    rec[c->rank] = mine;
    comm_ok = all_gather(rec.data());
    for (int p = 0; p < c->nranks && comm_ok; p++) all = all && rec[p].ok != 0;
  } else {
    c->peer_base[c->rank] = own;
  }
  cudaFree(dx);
  if (!(comm_ok && all)) {
    peer_release(c);
    c->peer_state = -1, c->peer_floats = frame_floats;
    return -1;
  }
  c->peer_state = 1, c->peer_floats = frame_floats;
  return 1;
}

// Delivers `floats` of a device-resident frame to the caller's `image` (device: enqueue-only; host: blocks).
int deliver_frame(mb200_comm *c, const float *d_frame, size_t floats, float *image) {
  mb200_scene *s = c->scene;
  if (device_pointer(image)) {
    if (cudaMemcpyAsync(image, d_frame, floats * sizeof(float), cudaMemcpyDeviceToDevice, s->stream) != cudaSuccess)
      return mb200::capi_set_error(MB200_ERR_CUDA, "gathered frame: device copy failed");
    return MB200_OK;
  }
  void *host = image;
  const bool pinned = pinned_pointer(image);
  if (!pinned) {
    if (c->pinned_cap < floats * sizeof(float)) {
      if (c->pinned) cudaFreeHost(c->pinned);
      c->pinned = nullptr, c->pinned_cap = 0;
      if (cudaMallocHost(&c->pinned, floats * sizeof(float)) != cudaSuccess)
        return mb200::capi_set_error(MB200_ERR_OUT_OF_MEMORY, "pinned staging for the gathered frame");
      c->pinned_cap = floats * sizeof(float);
    }
    host = c->pinned;
  }
  if (cudaMemcpyAsync(host, d_frame, floats * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
      cudaStreamSynchronize(s->stream) != cudaSuccess)
    return mb200::capi_set_error(MB200_ERR_CUDA, "gathered frame: device -> host copy failed");
  if (!pinned) memcpy(image, c->pinned, floats * sizeof(float));
  return MB200_OK;
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
int mb200_comm_exchange_path(const mb200_comm *c) { return c ? c->last_path : -1; }
int mb200_comm_rank(const mb200_comm *c) { return c ? c->rank : -1; }

void mb200_comm_destroy(mb200_comm *c) {
  if (!c) return;
  if (c->scene) {
    cudaSetDevice(c->scene->device);
    cudaStreamSynchronize(c->scene->stream);
  }
  peer_release(c);
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
  if (row_floats % 4 == 0 && peer_prepare(c, full_floats) == 1) { // (the same decision on every rank)
    // exchange over peer memory: the rows go straight into place in every rank's frame, frames alternate by epoch so
    // that a fast rank's next frame never lands in the buffer a slow rank is still reading
    if (local_rows > 0 && (reinterpret_cast<size_t>(d_bands) & 15u) != 0) { // float4 loads: through the aligned send buffer
      if ((rc = grow((void **)&c->send, &c->send_cap, send_floats * sizeof(float), s->stream)) != MB200_OK) return rc;
      if (cudaMemcpyAsync(c->send, d_bands, (size_t)local_rows * row_floats * sizeof(float), cudaMemcpyDeviceToDevice, s->stream) != cudaSuccess)
        return mb200::capi_set_error(MB200_ERR_CUDA, "band copy failed");
      d_bands = c->send;
    }
    c->epoch++;
    PeerTable t;
    memset(&t, 0, sizeof(t));
    for (int p = 0; p < c->nranks; p++) t.base[p] = c->peer_base[p];
    const size_t frame_off = kPeerHeaderBytes + (size_t)(c->epoch & 1ull) * full_floats * sizeof(float);
    const size_t vec = (size_t)local_rows * (row_floats / 4);
    size_t ctas = (vec + 255) / 256;
    if (ctas > 148 * 4) ctas = 148 * 4;
    if (ctas < 1) ctas = 1;
    k_exchange_rows<<<(unsigned)ctas, 256, 0, s->stream>>>((const float4 *)d_bands, t, frame_off, local_rows, (int)(row_floats / 4),
                                                          band_rows, c->nranks, c->rank, c->epoch);
    mb200::note_launch();
    if (cudaGetLastError() != cudaSuccess) return mb200::capi_set_error(MB200_ERR_CUDA, "row exchange launch failed");
    c->last_path = 1;
    if (!image) return MB200_OK;
    return deliver_frame(c, reinterpret_cast<const float *>(c->peer_base[c->rank] + frame_off), full_floats, image);
  }
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
  c->last_path = 0;
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

// fp64_halfwarp_probe.cu -- does the FP64 pipe of sm_100 charge a warp instruction by its ACTIVE lanes?
// A DFMA chain is run by a lane subset given as a mask: all 32 lanes, the lower 16, 8 + 8 spread over both halves,
// every other lane, 8 lanes of one half, 1 lane.  If the time per warp instruction followed the occupied half-warps (or
// quarter-warps), packing the lanes of one traversal phase into one half would relieve the pipe.  Development probe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_halfwarp_probe fp64_halfwarp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double *out, int iters, double a, double b, unsigned mask) {
  const unsigned lane = threadIdx.x & 31u;
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  if ((mask >> lane) & 1u) {
    for (int i = 0; i < iters; i++) {
      x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a);
      x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 20000;
  double *d;
  cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const struct { const char *name; unsigned mask; } cases[] = {
      {"all 32 lanes", 0xFFFFFFFFu}, {"lower 16 lanes", 0x0000FFFFu}, {"8 + 8 lanes (both halves)", 0x00FF00FFu},
      {"every other lane (16)", 0x55555555u}, {"8 lanes of one half", 0x000000FFu}, {"8 lanes, 2 per quarter", 0x03030303u},
      {"1 lane", 0x00000001u}};
  for (const auto &c : cases)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k<<<blocks, threads>>>(d, iters, 1e-3, 1.0000001, c.mask);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double winst = (double)blocks * (threads / 32) * iters * 8;
      if (rep)
        printf("%-28s %8.3f ms  %.2f cycles per warp-DFMA per SM sub-partition (at %d MHz)\n", c.name, ms,
               ms * 1e-3 * pr.clockRate * 1e3 / (winst / (pr.multiProcessorCount * 4)), pr.clockRate / 1000);
    }
  return 0;
}

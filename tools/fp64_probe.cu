// fp64_probe.cu -- measures B200 FP64 (DADD / DMUL / DFMA) and FP32 issue rates and L2-resident
// gather latency.  Not part of the product: numbers feed DESIGN.md's roofline discussion.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __global__ void k(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    if (OP == 0) { x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a); x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a); }
    if (OP == 1) { x0 = __dmul_rn(x0, b); x1 = __dmul_rn(x1, b); x2 = __dmul_rn(x2, b); x3 = __dmul_rn(x3, b); x4 = __dmul_rn(x4, b); x5 = __dmul_rn(x5, b); x6 = __dmul_rn(x6, b); x7 = __dmul_rn(x7, b); }
    if (OP == 2) { x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a); x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void kf(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-9f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) { x0 = __fmaf_rn(x0, b, a); x1 = __fmaf_rn(x1, b, a); x2 = __fmaf_rn(x2, b, a); x3 = __fmaf_rn(x3, b, a); x4 = __fmaf_rn(x4, b, a); x5 = __fmaf_rn(x5, b, a); x6 = __fmaf_rn(x6, b, a); x7 = __fmaf_rn(x7, b, a); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void kdiv(double *out, int iters, double a) {
  double x0 = 1.0 + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  for (int i = 0; i < iters; i++) { x0 = a / x0; x1 = a / x1; x2 = a / x2; x3 = a / x3; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
// dependent pointer chase through an L2-resident table of 128-byte lines
__global__ void kchase(const unsigned *next, unsigned *out, int iters) {
  unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) * 32u % (1u << 20);
  for (int i = 0; i < iters; i++) p = next[p];
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs=%d clock=%d kHz L2=%d MB\n", pr.name, pr.multiProcessorCount, pr.clockRate, pr.l2CacheSize >> 20);
  const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 20000;
  double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char *names[3] = {"DADD", "DMUL", "DFMA"};
  for (int op = 0; op < 3; op++) for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    if (op == 0) k<0><<<blocks, threads>>>(d, iters, 1e-3, 1.0000001);
    if (op == 1) k<1><<<blocks, threads>>>(d, iters, 1e-3, 1.0000001);
    if (op == 2) k<2><<<blocks, threads>>>(d, iters, 1e-3, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8;
    if (rep) printf("%s: %.2f Tops/s (%.1f ops/clk/SM at %d MHz nominal)\n", names[op], ops / ms * 1e-9, ops / (ms * 1e-3) / pr.multiProcessorCount / (pr.clockRate * 1e3), pr.clockRate / 1000);
  }
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0); kf<<<blocks, threads>>>((float *)d, iters, 1e-3f, 1.0000001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); double ops = (double)blocks * threads * iters * 8;
    if (rep) printf("FFMA: %.2f Tops/s (%.1f ops/clk/SM)\n", ops / ms * 1e-9, ops / (ms * 1e-3) / pr.multiProcessorCount / (pr.clockRate * 1e3));
  }
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0); kdiv<<<blocks, threads>>>(d, 2000, 3.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); double ops = (double)blocks * threads * 2000 * 4;
    if (rep) printf("DDIV: %.3f Tops/s (%.2f div/clk/SM)\n", ops / ms * 1e-9, ops / (ms * 1e-3) / pr.multiProcessorCount / (pr.clockRate * 1e3));
  }
  { // pointer chase: 1M entries * 4B = 4 MB table, stride pattern jumps 128B lines pseudo-randomly
    const unsigned N = 1u << 20; unsigned *h = new unsigned[N];
    for (unsigned i = 0; i < N; i++) h[i] = (unsigned)(((unsigned long long)i * 1664525ull + 1013904223ull) % N);
    unsigned *dn, *dout; cudaMalloc(&dn, N * 4); cudaMalloc(&dout, 4 * 1024); cudaMemcpy(dn, h, N * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0); kchase<<<1, 32>>>(dn, dout, 20000); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("L2-resident dependent gather: %.0f ns per hop (one warp, 32 distinct lines)\n", ms * 1e6 / 20000);
    }
  }
  return 0;
}

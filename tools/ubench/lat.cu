// Microbenchmark: dependent-issue latency of FP64 ops on sm_100a, and throughput with k independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dfma_chain(double* out, int iters, double a, double b, long long* cyc) {
  double x[CH];
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcp_chain(double* out, int iters, long long* cyc) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      double y;
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = y + 1.0;
    }
  }
  long long t1 = clock64();
  if (x == 1234.5) out[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void div_chain(double* out, int iters, double a, long long* cyc) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x = a / x + 1.0;
  }
  long long t1 = clock64();
  if (x == 1234.5) out[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  long long h;
  const int it = 2000;
#define RUN(name, kern, warps, nops)                                               \
  kern; cudaDeviceSynchronize(); kern; cudaDeviceSynchronize();                  \
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);                                  \
  printf("%-40s warps/block %2d : %.2f cycles per op-slot (per warp)\n", name, warps, (double)h / (it * 16.0 * nops));
  for (int w : {1, 2, 4, 8, 16}) {
    RUN("DFMA 1 chain", (dfma_chain<1><<<1, 32 * w>>>(d, it, 1.0000001, 1e-9, c)), w, 1)
    RUN("DFMA 2 chains", (dfma_chain<2><<<1, 32 * w>>>(d, it, 1.0000001, 1e-9, c)), w, 2)
    RUN("DFMA 4 chains", (dfma_chain<4><<<1, 32 * w>>>(d, it, 1.0000001, 1e-9, c)), w, 4)
    RUN("DFMA 8 chains", (dfma_chain<8><<<1, 32 * w>>>(d, it, 1.0000001, 1e-9, c)), w, 8)
  }
  for (int w : {1, 4}) {
    RUN("rcp.approx + DADD chain (2 ops)", (rcp_chain<<<1, 32 * w>>>(d, it, c)), w, 1)
    RUN("IEEE div + DADD chain", (div_chain<<<1, 32 * w>>>(d, it, 3.0, c)), w, 1)
  }
  return 0;
}

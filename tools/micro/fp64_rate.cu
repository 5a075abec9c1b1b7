// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_rate tools/micro/fp64_rate.cu
// Micro-benchmark: per-SM throughput of DFMA, F2F.F64.F32 and FFMA on this GPU (run under gpurun).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, float seed, int iters) {
  double a[8]; float f[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; f[i] = seed + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(a[i], 1.0000001, 0.5);
      if (MODE == 1) { a[i] += (double)f[i]; f[i] = __int_as_float(__float_as_int(f[i]) + 1); }
      if (MODE == 2) f[i] = fmaf(f[i], 1.0000001f, 0.5f);
      if (MODE == 3) { a[i] = fma((double)f[i], 1.5, a[i]); f[i] = __int_as_float(__float_as_int(f[i]) + 1); }
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double ops_per_iter) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(out, 1.0f, 100);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, 1.0f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = (double)148 * 8 * 256 * iters * 8 * ops_per_iter;
  printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at 1.965 GHz)\n", name, ms, total / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 1); run<1>("F2F.F64.F32 + DADD", 2); run<2>("FFMA", 1); run<3>("F2F + DFMA", 2);
  return 0;
}

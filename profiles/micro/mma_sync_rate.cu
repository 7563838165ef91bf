// Microbenchmark: legacy warp-level mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16), FMA/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tf32(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x, 7};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x, 7};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16}) {
    for (int which = 0; which < 2; ++which) {
      const int iters = 20000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_tf32<<<148, warps * 32>>>(out, iters); else k_bf16<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma_per_mma = which == 0 ? 16.0 * 8 * 8 : 16.0 * 8 * 16;
      const double fma = (double)iters * 8 * warps * fma_per_mma;   // per SM
      printf("%s warps/SM=%2d  %.3f ms  %.0f FMA/clk/SM (at %d MHz nominal)  %.1f TFLOP/s chip\n", which == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16",
             warps, ms, fma / (ms * 1e-3) / (clk * 1e3), clk / 1000, 2 * fma * 148 / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}

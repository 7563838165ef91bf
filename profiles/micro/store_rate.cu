// Store-path micro benchmark: how fast can 148 CTAs write a [M, N] fp32 matrix in the access pattern of the GEMM
// epilogue (CTA tile 128 rows x 256 columns, warp = 32 rows x 64 columns), by store mechanism:
//   0  TMA tensor store, 32 x 32 fp32 boxes (128-byte rows), SWIZZLE_128B   (what gemm_tc*.cu does)
//   1  cp.async.bulk 1-D, one 256-byte row piece per copy (32 per warp region)
//   2  st.global.v4, lane = row (each lane walks its own 128-byte row piece)
//   3  st.global.v4, lanes along the row (512 contiguous bytes per warp instruction)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_rate store_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(const __grid_constant__ CUtensorMap tmC, float* C, int M, int N) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lg = warp & 3, cq = warp >> 2;                 // 32-row group, 64-column quarter of the 128 x 256 CTA tile
  unsigned char* box = base + warp * (MODE == 5 ? 32768 : 8192);   // two 4 KB boxes per warp (mode 5: 32 KB, warps 0-3 only)
  for (int i = lane; i < (MODE == 5 ? 8192 : 2048); i += 32) reinterpret_cast<float*>(box)[i] = (float)(i + warp);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int m_tiles = (M + 127) / 128, n_tiles = N / 256;
  int sbuf = 0, pending = 0;
  for (int t = blockIdx.x; t < m_tiles * n_tiles; t += gridDim.x) {
    const int nt = t % n_tiles, mt = t / n_tiles;
    const int row0 = mt * 128 + lg * 32;
    if (row0 >= M) continue;
    if (MODE == 5) {
      if (warp < 4) {
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        // (a real epilogue would refill the staging rows here)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        float* g = C + (size_t)(row0 + lane) * N + nt * 256;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1024;" ::"l"(g), "r"(smem_u32(box + lane * 1024)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      continue;
    }
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      const int col0 = nt * 256 + cq * 64 + ch * 32;
      unsigned char* b = box + sbuf * 4096;
      if (MODE == 0 || MODE == 4) {
        if (pending >= 2) {
          if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        *reinterpret_cast<float4*>(b + lane * 128 + ((0 ^ (lane & 7)) << 4)) = make_float4(1.f, 2.f, 3.f, (float)t);   // touch the box
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                       "r"(smem_u32(b)), "r"(col0), "r"(row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        sbuf ^= 1;
        if (pending < 2) ++pending;
      } else if (MODE == 1) {
        if (ch == 0) {
          if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          float* g = C + (size_t)(row0 + lane) * N + nt * 256 + cq * 64;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(g), "r"(smem_u32(box + lane * 256)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (MODE == 2) {
        float4* g = reinterpret_cast<float4*>(C + (size_t)(row0 + lane) * N + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = make_float4(1.f, 2.f, (float)j, (float)t);
      } else if (MODE == 3) {
        if (ch == 0) {
          // 32 rows x 64 columns = 32 rows x 16 float4: half a warp per row, two rows per instruction
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            float4* g = reinterpret_cast<float4*>(C + (size_t)(row0 + r + (lane >> 4)) * N + nt * 256 + cq * 64) + (lane & 15);
            *g = make_float4(1.f, 2.f, (float)r, (float)t);
          }
        }
      }
    }
  }
  if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5) {
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <int MODE>
void run(const CUtensorMap& tm, float* C, int M, int N, const char* name) {
  const size_t smem = (MODE == 5 ? 4 * 32768 : 16 * 8192) + 1024;
  CK(cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) store_kernel<MODE><<<148, 512, smem>>>(tm, C, M, N);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int i = 0; i < reps; ++i) store_kernel<MODE><<<148, 512, smem>>>(tm, C, M, N);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1000.0 / reps;
  printf("N=%4d mode %d %-46s %7.1f us  %6.0f GB/s\n", N, MODE, name, us, (double)M * N * 4 / us / 1e3);
}

int main() {
  const int M = 2048 * 65;
  for (int N : {1024, 768, 256}) {
    float* C;
    CK(cudaMalloc(&C, (size_t)M * N * 4));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map: %d\n", (int)r); return 1; }
    run<0>(tm, C, M, N, "TMA tensor store 32x32 fp32 (128 B rows)");
    run<1>(tm, C, M, N, "bulk 1-D, 256 B row pieces");
    run<2>(tm, C, M, N, "st.global.v4, lane = row");
    run<3>(tm, C, M, N, "st.global.v4, lanes along the row");
    CK(cudaFree(C));
  }
  return 0;
}

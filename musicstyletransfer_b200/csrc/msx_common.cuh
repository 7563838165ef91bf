// Shared helpers for libmsx.so (sm_100a only).  Host side: error reporting across the C ABI.
// Device side: warp primitives, Philox4x32-10 for dropout / sampling, vector load/store helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MSX_OK 0
#define MSX_ERR_ARG (-1)
#define MSX_ERR_CUDA (-2)
#define MSX_ERR_UNSUPPORTED (-3)

void msx_set_error(const char* fmt, ...);

#define MSX_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      msx_set_error(__VA_ARGS__);              \
      return MSX_ERR_ARG;                      \
    }                                          \
  } while (0)

#define MSX_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      msx_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                    cudaGetErrorString(_e));                                        \
      return MSX_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define MSX_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                         \
    if (_e != cudaSuccess) {                                                        \
      msx_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,        \
                    cudaGetErrorString(_e));                                        \
      return MSX_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

int msx_num_sms();

static inline int msx_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
#define MSX_FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MSX_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(MSX_FULL, v, o));
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: (seed, subsequence/offset) -> 4 x u32.
struct Philox {
  static constexpr uint32_t kA = 0xD2511F53u, kB = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __device__ __forceinline__ static uint4 gen(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(kA, c0), lo0 = kA * c0;
      uint32_t hi1 = __umulhi(kB, c2), lo1 = kB * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += kW0; k1 += kW1;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

// Dropout keep-mask for element `idx` of dropout site `site`: one Philox block yields 4 decisions.
// keep iff u >= p.  Forward and backward regenerate the same mask from (seed, site, idx).
__device__ __forceinline__ float dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx4, float p, float inv_keep,
                                                float out[4]) {
  uint4 r = Philox::gen(seed, idx4, (uint64_t)site);
  out[0] = u32_to_unit(r.x) >= p ? inv_keep : 0.f;
  out[1] = u32_to_unit(r.y) >= p ? inv_keep : 0.f;
  out[2] = u32_to_unit(r.z) >= p ? inv_keep : 0.f;
  out[3] = u32_to_unit(r.w) >= p ? inv_keep : 0.f;
  return 0.f;
}
#endif

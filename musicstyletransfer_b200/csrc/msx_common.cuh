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
// Optional device-side step counter added to every dropout / eps seed (msx_set_step_counter): lets a CUDA graph of the
// train step draw fresh masks on every replay although the host-side seed argument is frozen in the graph.
const unsigned long long* msx_step_counter();

static inline int msx_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (msx_set_pdl / MSX_PDL, default on).  A kernel that calls pdl_entry() as its first
// statement may be launched through msx_launch(): its grid is then scheduled while the previous kernel of the stream is
// still running (CTA launch, parameter and constant loads overlap the predecessor's tail) and blocks in
// griddepcontrol.wait until every prerequisite grid has completed and flushed, so the stream's memory ordering is
// unchanged.  A step is a chain of 60-75 dependent launches; inside a CUDA graph the hand-over between two kernels
// otherwise costs 2-3 us each.  Kernels launched with <<< >>> keep the full serialisation (griddepcontrol is a no-op there).
int msx_pdl_enabled();

#ifdef __CUDACC__
#define MSX_FULL 0xffffffffu

// First statement of every kernel launched through msx_launch(): let the next kernel of the stream be scheduled, then wait
// for the previous one(s).  Nothing before the wait may touch global memory or allocate tensor memory (a dependent CTA that
// held TMEM columns while a predecessor CTA on the same SM still waits for its own would deadlock).
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t msx_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = msx_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}

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

// Transpose-reduce: every lane holds v[0..31]; afterwards lane L holds sum over all lanes of v[L] (in v[0]).
// 31 shuffles instead of 32 x 5; used to fold bias-gradient column sums into row-owner epilogues.
__device__ __forceinline__ float warp_colsum32(float v[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(MSX_FULL, send, s);
      v[i] = (up ? v[i + s] : v[i]) + recv;
    }
  }
  return v[0];
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
__device__ __forceinline__ unsigned long long msx_eff_seed(unsigned long long seed, const unsigned long long* ctr) {
  return ctr ? seed + __ldg(ctr) : seed;
}
// bf16 hi / lo planes (operands of the "p3" forward GEMMs: hi*hi + hi*lo + lo*hi on kind::f16, ~2^-17 per product):
// hi = rn_bf16(x), lo = rn_bf16(x - hi), four values -> two packed 8-byte words (element 0 in the low half of .x)
__device__ __forceinline__ void split4_bf16(float a, float b, float c, float d, uint2& hi, uint2& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.x) : "f"(b), "f"(a));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.y) : "f"(d), "f"(c));
  const float ra = a - __uint_as_float(hi.x << 16), rb = b - __uint_as_float(hi.x & 0xFFFF0000u);
  const float rc = c - __uint_as_float(hi.y << 16), rd = d - __uint_as_float(hi.y & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo.x) : "f"(rb), "f"(ra));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo.y) : "f"(rd), "f"(rc));
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

// Dropout keep-mask for elements 4*idx4 .. 4*idx4+3 of dropout site `site`: keep iff u >= p.  Forward and
// backward regenerate the same mask from (seed, site, element index), nothing is stored.  The generator is a
// counter-based hash (a Weyl multiply plus the murmur3 32-bit finaliser over the element-pair index, keyed by a
// 32-bit key derived from seed and site); each 32-bit hash serves TWO elements as 16-bit uniforms (keep iff
// u16 >= floor(p * 65536), i.e. p is honoured to 1.5e-5), ~7 integer instructions per element instead of ~25 for
// Philox4x32-10, which made the GEMM epilogue of the FF1 layer RNG-bound.  Philox stays in use for eps ~ N(0,1) and
// token sampling.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx4, float p, float inv_keep,
                                                float out[4]) {
  const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (site + 1u)));
  const uint32_t hi = mix32(key ^ (uint32_t)(idx4 >> 31));
  const uint32_t base = (uint32_t)idx4 << 1;
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const uint32_t r0 = mix32((base ^ hi) * 0x9E3779B1u + key);
  const uint32_t r1 = mix32(((base + 1u) ^ hi) * 0x9E3779B1u + key);
  out[0] = (r0 & 0xFFFFu) >= thr ? inv_keep : 0.f;
  out[1] = (r0 >> 16) >= thr ? inv_keep : 0.f;
  out[2] = (r1 & 0xFFFFu) >= thr ? inv_keep : 0.f;
  out[3] = (r1 >> 16) >= thr ? inv_keep : 0.f;
  return 0.f;
}
// Same masks as dropout_scale4 for the 32 consecutive elements e0 .. e0+31 of one row chunk, applied in place
// (v[j] = keep ? v[j] * inv_keep : 0): the key and (outside the one chunk in 2^33 elements that straddles it) the
// high-word hash are computed once per chunk, and the 16-bit threshold tests are done on the full hash word.
__device__ __forceinline__ void dropout_apply32(uint64_t seed, uint32_t site, uint64_t e0, float p, float inv_keep,
                                                float v[32]) {
  const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (site + 1u)));
  const uint64_t idx4_0 = e0 >> 2;
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const uint32_t thr_hi = thr << 16;
  const uint32_t hi0 = mix32(key ^ (uint32_t)(idx4_0 >> 31));
  const bool same_hi = ((idx4_0 + 7) >> 31) == (idx4_0 >> 31);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint64_t idx4 = idx4_0 + g;
    const uint32_t hi = same_hi ? hi0 : mix32(key ^ (uint32_t)(idx4 >> 31));
    const uint32_t base = (uint32_t)idx4 << 1;
    const uint32_t r0 = mix32((base ^ hi) * 0x9E3779B1u + key);
    const uint32_t r1 = mix32(((base + 1u) ^ hi) * 0x9E3779B1u + key);
    v[4 * g + 0] = ((r0 << 16) >= thr_hi) ? v[4 * g + 0] * inv_keep : 0.f;
    v[4 * g + 1] = (r0 >= thr_hi) ? v[4 * g + 1] * inv_keep : 0.f;
    v[4 * g + 2] = ((r1 << 16) >= thr_hi) ? v[4 * g + 2] * inv_keep : 0.f;
    v[4 * g + 3] = (r1 >= thr_hi) ? v[4 * g + 3] * inv_keep : 0.f;
  }
}
#endif

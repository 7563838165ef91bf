// libmsx.so core: version, thread-local error string, device properties cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "msx_common.cuh"

static thread_local char g_last_error[512] = "";

void msx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int msx_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" int msx_version(void) { return 100; }  // 0.1.0

// ---- programmatic dependent launch switch (see msx_common.cuh): MSX_PDL=0 in the environment or msx_set_pdl(0) turn it off
static int g_pdl = -1;
int msx_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("MSX_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl;
}
extern "C" int msx_set_pdl(int on) {
  g_pdl = on ? 1 : 0;
  return MSX_OK;
}
extern "C" int msx_get_pdl(void) { return msx_pdl_enabled(); }

extern "C" const char* msx_last_error(void) { return g_last_error; }

extern "C" int msx_device_sm_count(void) { return msx_num_sms(); }

// ---- device-side step counter for graph-replayed steps (see msx_common.cuh)
static unsigned long long* g_step_counter = nullptr;
const unsigned long long* msx_step_counter() { return g_step_counter; }

extern "C" int msx_set_step_counter(unsigned long long* dev_counter) {
  g_step_counter = dev_counter;
  return MSX_OK;
}

static __global__ void step_counter_tick_kernel(unsigned long long* c) {
  pdl_entry(); *c += 1ull; }

extern "C" int msx_step_counter_tick(unsigned long long* dev_counter, void* stream) {
  MSX_REQUIRE(dev_counter, "msx_step_counter_tick: null pointer");
  MSX_CUDA(msx_launch(step_counter_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, dev_counter));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

// ---- fp32 -> bf16 cast (operands of the bf16 GEMM variant)
static __global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, unsigned* __restrict__ dst,
                                                                    long long n) {
  pdl_entry();
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
      uint4 o;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(a.y), "f"(a.x));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(a.w), "f"(a.z));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(b.y), "f"(b.x));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(b.w), "f"(b.z));
      *reinterpret_cast<uint4*>(dst + i / 2) = o;
    } else {
      unsigned short* d16 = reinterpret_cast<unsigned short*>(dst);
      for (long long j = i; j < n; ++j) {
        unsigned r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(src[j]));
        d16[j] = (unsigned short)(r & 0xFFFFu);
      }
    }
  }
}

extern "C" int msx_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  MSX_REQUIRE(n >= 0, "msx_cast_f32_bf16: negative length");
  if (n == 0) return MSX_OK;
  MSX_REQUIRE(src && dst, "msx_cast_f32_bf16: null pointer");
  MSX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "msx_cast_f32_bf16: pointers must be 16-byte aligned");
  const long long want = (n + 8 * 256 - 1) / (8 * 256);
  const int grid = (int)(want < (long long)msx_num_sms() * 8 ? want : (long long)msx_num_sms() * 8);
  MSX_CUDA(msx_launch(cast_f32_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src, reinterpret_cast<unsigned*>(dst), n));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

// ---- keep-mask of one dropout site, as the step's kernels draw it (msx_common.cuh: dropout_scale4): out[e] = 1 when
// element e (row-major index into the site's activation matrix) is kept.  Lets a checker replay a dropout step.
static __global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ out, long long n, float p,
                                                                   unsigned long long seed, const unsigned long long* ctr,
                                                                   unsigned site) {
  pdl_entry();
  const unsigned long long eff = msx_eff_seed(seed, ctr);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += stride) {
    float k[4];
    dropout_scale4(eff, site, (uint64_t)q, p, 1.f, k);
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = k[j] != 0.f ? 1 : 0;
  }
}

extern "C" int msx_dropout_mask(uint8_t* out, long long n, float drop_p, unsigned long long seed, unsigned site, void* stream) {
  MSX_REQUIRE(n >= 0, "msx_dropout_mask: negative length");
  if (n == 0) return MSX_OK;
  MSX_REQUIRE(out, "msx_dropout_mask: null pointer");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_dropout_mask: dropout probability must be in [0,1)");
  const long long want = (n / 4 + 256) / 256;
  const int grid = (int)(want < (long long)msx_num_sms() * 8 ? want : (long long)msx_num_sms() * 8);
  MSX_CUDA(msx_launch(dropout_mask_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, out, n, drop_p, seed, msx_step_counter(), site));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

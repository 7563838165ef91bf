// libmsx.so core: version, thread-local error string, device properties cache.
#include <stdarg.h>
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

extern "C" const char* msx_last_error(void) { return g_last_error; }

extern "C" int msx_device_sm_count(void) { return msx_num_sms(); }

// ---- device-side step counter for graph-replayed steps (see msx_common.cuh)
static unsigned long long* g_step_counter = nullptr;
const unsigned long long* msx_step_counter() { return g_step_counter; }

extern "C" int msx_set_step_counter(unsigned long long* dev_counter) {
  g_step_counter = dev_counter;
  return MSX_OK;
}

static __global__ void step_counter_tick_kernel(unsigned long long* c) { *c += 1ull; }

extern "C" int msx_step_counter_tick(unsigned long long* dev_counter, void* stream) {
  MSX_REQUIRE(dev_counter, "msx_step_counter_tick: null pointer");
  step_counter_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_counter);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

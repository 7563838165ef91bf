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

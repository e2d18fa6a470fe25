// Error plumbing and device queries for the C ABI (include/cmlpl.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cmlpl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
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

}  // namespace cmlpl

extern "C" {

int cmlpl_version(void) { return 100; }

const char* cmlpl_last_error(void) { return cmlpl::g_err; }

int cmlpl_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cmlpl::set_error("cmlpl_device_ok: no CUDA device");
    return CMLPL_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

}  // extern "C"

// Error reporting, version and device query of the C ABI (include/lshm.h).
#include <stdarg.h>
#include <string>
#include "common.cuh"

namespace lshm {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace lshm

extern "C" {

const char* lshm_last_error(void) { return lshm::g_last_error.c_str(); }

int lshm_version(void) { return 100; }

int lshm_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  LSHM_CUDA(cudaGetDevice(&dev), "lshm_device_info");
  int a = 0, b = 0, c = 0;
  LSHM_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev), "lshm_device_info");
  LSHM_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev), "lshm_device_info");
  LSHM_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev), "lshm_device_info");
  if (sm) *sm = a;
  if (cc_major) *cc_major = b;
  if (cc_minor) *cc_minor = c;
  return LSHM_OK;
}

}  // extern "C"

// api.cu -- error plumbing, version and device queries of the C ABI.
#include "common.cuh"

#include <string.h>

namespace cmr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_cuda(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return CMR_ECUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no CUDA device: libcmrag has no CPU fallback");
    return -1;
  }
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      set_error("cannot query SM count");
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace cmr

extern "C" const char* cmr_last_error(void) { return cmr::g_err; }

extern "C" int cmr_version(void) { return 100; }

extern "C" int cmr_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  CMR_CUDA(cudaGetDevice(&dev));
  int n = cmr::sm_count();
  if (n <= 0) return CMR_ECUDA;
  int maj = 0, min = 0;
  CMR_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  CMR_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_out) *sm_count_out = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return CMR_OK;
}

// gtok_api.cu - version, thread-local error string, device queries.
#include <stdarg.h>
#include <string.h>

#include "gtok_common.cuh"

namespace gtok {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return GTOK_ECUDA;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}

}  // namespace gtok

extern "C" int gtok_version(void) { return GTOK_VERSION; }

extern "C" const char* gtok_last_error(void) { return gtok::g_err; }

extern "C" int gtok_device_sm_count(void) {
  int n = gtok::sm_count();
  if (n < 0) {
    gtok::set_error("no CUDA device available (gtok has no CPU fallback)");
    return GTOK_ECUDA;
  }
  return n;
}

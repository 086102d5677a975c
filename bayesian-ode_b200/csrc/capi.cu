// C-ABI plumbing: version, thread-local error string, device query.
#include "common.cuh"
#include <string.h>

namespace bode {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BODE_OK;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return BODE_ERR_CUDA;
}
}  // namespace bode

extern "C" int bode_version(void) { return 100; }

extern "C" const char* bode_last_error(void) { return bode::g_err; }

extern "C" int bode_device_sm_count(void) {
  int dev = 0, sms = 0;
  BODE_CUDA(cudaGetDevice(&dev));
  BODE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return sms;
}

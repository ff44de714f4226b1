// Error plumbing and library-level entry points of libwbk.
#include "wbk_common.cuh"
#include <cstdarg>

static thread_local char g_err[512] = "";

void wbk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* wbk_last_error(void) { return g_err; }

extern "C" int wbk_version(void) { return 100; }

extern "C" int wbk_device_count(void) {
#ifdef WBK_EMU
  return 1;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    wbk_set_error("no CUDA device available (%s); libwbk has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
    return WBK_ERR_NODEVICE;
  }
  return n;
#endif
}

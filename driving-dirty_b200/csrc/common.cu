// Error text, launch accounting, version.
#include <stdarg.h>

#include "dd_common.cuh"

namespace dd {
thread_local char g_err[512] = "";
long long g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}
}  // namespace dd

extern "C" int dd_version(void) { return 100; }

extern "C" int dd_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return static_cast<int>(strlen(dd::g_err));
  strncpy(buf, dd::g_err, len - 1);
  buf[len - 1] = 0;
  return static_cast<int>(strlen(buf));
}

extern "C" long long dd_launch_count(void) { return __atomic_load_n(&dd::g_launches, __ATOMIC_RELAXED); }

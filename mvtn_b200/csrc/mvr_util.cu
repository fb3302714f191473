// mvr_util.cu -- error reporting shared by the C-ABI entry points.
#include <cstdarg>
#include <cstdio>

#include "mvr_common.cuh"

namespace mvr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Launch-time errors only (bad configuration, missing kernel image): never synchronises.
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

}  // namespace mvr

extern "C" int mvr_abi_version(void) { return MVR_ABI_VERSION; }
extern "C" const char* mvr_last_error_string(void) { return mvr::g_err; }

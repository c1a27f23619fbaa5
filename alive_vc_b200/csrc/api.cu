// Error plumbing and version of the alive_knn C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace alive {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace alive

extern "C" const char* alive_knn_last_error(void) { return alive::g_error; }
extern "C" int alive_knn_abi_version(void) { return ALIVE_KNN_ABI_VERSION; }

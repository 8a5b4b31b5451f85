// Library identification and the thread-local error string of the C ABI (include/tn_b200.h).
#include <stdarg.h>

#include "tn_common.cuh"

namespace tn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tn

extern "C" int tn_version(void) { return 100; }
extern "C" const char* tn_last_error_string(void) { return tn::g_err; }
extern "C" const char* tn_build_arch(void) { return "sm_100a"; }

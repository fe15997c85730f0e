// common.cu -- error string + version entry points of the C ABI.
#include <stdarg.h>

#include "nsv_common.cuh"

namespace nsv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace nsv

extern "C" int nsv_version(void) { return 1; }
extern "C" const char* nsv_last_error_string(void) { return nsv::g_err; }
extern "C" const char* nsv_build_arch(void) { return "sm_100a"; }

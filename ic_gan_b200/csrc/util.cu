// Error reporting and version for the C ABI (thread-local last-error string; SURVEY.md §8b "Errors").
#include <stdarg.h>

#include "common.cuh"

namespace icgan {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace icgan

extern "C" const char* icgan_last_error(void) { return icgan::g_err; }
extern "C" int icgan_version(void) { return 100; }

// Error reporting / bookkeeping shared by every entry point of the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace fd {
static thread_local char g_err[512] = "";
long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace fd

extern "C" {
const char* fd_last_error(void) { return fd::g_err; }
int fd_version(void) { return 100; }
long fd_launch_count(void) { return fd::g_launches; }
}

// Library-wide plumbing of libdirb200: error string, version, launch counter.
#include "common.cuh"
#include <stdarg.h>

namespace dirb200 {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dirb200

extern "C" {
const char* dirb200_last_error(void) { return dirb200::g_err; }
int dirb200_version(void) { return 100; }
int64_t dirb200_launch_count(void) { return dirb200::g_launches.load(); }
}

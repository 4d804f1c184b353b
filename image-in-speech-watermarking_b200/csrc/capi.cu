// Library-wide pieces of the C ABI: version, thread-local error string, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "wmk_common.cuh"

namespace wmk {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace wmk

extern "C" int wmk_version(void) { return 100; }
extern "C" const char* wmk_last_error(void) { return wmk::g_err; }
extern "C" uint64_t wmk_launch_count(void) { return wmk::g_launches.load(std::memory_order_relaxed); }

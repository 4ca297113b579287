// Library-level bookkeeping of the C ABI: version, last error, launch counter.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace ac {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ac

extern "C" int ac_abi_version(void) { return AC_ABI_VERSION; }
extern "C" const char* ac_last_error(void) { return ac::g_err; }
extern "C" int64_t ac_launch_count(void) { return (int64_t)ac::g_launches.load(); }

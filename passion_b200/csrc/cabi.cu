// Library-level C-ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void pb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void pb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int pb_version(void) { return 100; }
extern "C" const char* pb_last_error(void) { return g_err; }
extern "C" long long pb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

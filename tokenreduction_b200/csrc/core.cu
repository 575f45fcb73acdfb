// Library-wide state: ABI version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace tokred {
static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace tokred

extern "C" int tokred_abi_version(void) { return TOKRED_ABI_VERSION; }
extern "C" const char* tokred_last_error(void) { return tokred::g_error; }
extern "C" uint64_t tokred_launch_count(void) { return tokred::g_launches.load(std::memory_order_relaxed); }

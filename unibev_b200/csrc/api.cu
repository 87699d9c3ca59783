// Library-wide state of libunibev_b200: thread-local error text, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "ub_common.cuh"

namespace ub {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
}  // namespace ub

extern "C" int ub_version(void) { return 1000; }
extern "C" const char* ub_last_error(void) { return ub::g_err; }
extern "C" int64_t ub_launch_count(void) { return ub::g_launches.load(std::memory_order_relaxed); }
extern "C" void ub_launch_count_reset(void) { ub::g_launches.store(0, std::memory_order_relaxed); }
// Programmatic dependent launch of the GEMM and sampling kernels (kernel prologues overlap the predecessor's tail): on by
// default; a performance knob, results do not depend on it.
extern "C" int ub_set_pdl(int on) {
  ub::g_pdl = on ? 1 : 0;
  return UB_OK;
}

// Library-wide state of libunibev_b200: thread-local error text, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>

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
static std::atomic<int64_t> g_unsupported{0};
int unsupported() {
  g_unsupported.fetch_add(1, std::memory_order_relaxed);
  return UB_EUNSUPPORTED;
}
namespace {
std::mutex g_dev_mu;
std::map<int, int> g_sm_count;                                   // device ordinal -> SMs
std::map<std::pair<const void*, int>, size_t> g_smem_configured;  // (kernel, device) -> opted-in dynamic shared memory
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
}  // namespace
int sm_count() {
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_dev_mu);
  auto it = g_sm_count.find(dev);
  if (it != g_sm_count.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = kNumSMs;
  }
  g_sm_count[dev] = n;
  return n;
}
bool smem_config_needed(const void* kernel, size_t smem) {
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_dev_mu);
  auto it = g_smem_configured.find({kernel, dev});
  return it == g_smem_configured.end() || it->second < smem;
}
void smem_config_done(const void* kernel, size_t smem) {
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_dev_mu);
  g_smem_configured[{kernel, dev}] = smem;
}
static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
}  // namespace ub

extern "C" int ub_version(void) { return 1000; }
extern "C" const char* ub_last_error(void) { return ub::g_err; }
extern "C" int64_t ub_launch_count(void) { return ub::g_launches.load(std::memory_order_relaxed); }
extern "C" void ub_launch_count_reset(void) {
  ub::g_launches.store(0, std::memory_order_relaxed);
  ub::g_unsupported.store(0, std::memory_order_relaxed);
}
// how many calls returned UB_EUNSUPPORTED since the last reset: every one of them made the caller take a generic
// (slower) entry point -- bench.py asserts the timed step has none
extern "C" int64_t ub_unsupported_count(void) { return ub::g_unsupported.load(std::memory_order_relaxed); }
// Programmatic dependent launch of the GEMM and sampling kernels (kernel prologues overlap the predecessor's tail): on by
// default; a performance knob, results do not depend on it.
extern "C" int ub_set_pdl(int on) {
  ub::g_pdl = on ? 1 : 0;
  return UB_OK;
}

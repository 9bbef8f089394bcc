// Library-wide plumbing: version, per-thread error string, device properties, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <cstdlib>

#include "common.cuh"

namespace l2s {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int query_attr(cudaDeviceAttr a, int fallback) {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fallback;
  if (cudaDeviceGetAttribute(&v, a, dev) != cudaSuccess) return fallback;
  return v;
}

bool pdl_enabled() {
  const char* e = getenv("L2S_PDL");
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  return false;     // measured on B200 inside the whole-step CUDA graph: 11.05 ms with, 10.57 ms without -> opt-in only
}

int sm_count() {
  static thread_local int cached = 0;
  if (!cached) cached = query_attr(cudaDevAttrMultiProcessorCount, 148);
  return cached;
}

int max_smem_optin() {
  static thread_local int cached = 0;
  if (!cached) cached = query_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, 232448);
  return cached;
}

}  // namespace l2s

extern "C" int l2s_version(void) { return 100; }
extern "C" const char* l2s_last_error_string(void) { return l2s::g_err; }
extern "C" uint64_t l2s_launch_count(void) { return l2s::g_launches.load(std::memory_order_relaxed); }

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

// device attributes are cached per device id (a process may drive several GPUs from one thread)
static int cached_attr(cudaDeviceAttr a, int which, int fallback) {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[2][kMaxDev];          // zero-initialised; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fallback;
  if (dev < 0 || dev >= kMaxDev) {
    int v = 0;
    return cudaDeviceGetAttribute(&v, a, dev) == cudaSuccess ? v : fallback;
  }
  int v = cache[which][dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, a, dev) != cudaSuccess || v <= 0) return fallback;
    cache[which][dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// environment switches are read ONCE (first use): no getenv on the launch path, no race with setenv afterwards
bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}

bool pdl_enabled() {
  // measured on B200 inside the whole-step CUDA graph: 11.05 ms with, 10.57 ms without -> opt-in only
  static const bool on = env_flag("L2S_PDL");
  return on;
}

int sm_count() { return cached_attr(cudaDevAttrMultiProcessorCount, 0, 148); }

int max_smem_optin() { return cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, 1, 232448); }

// diagnostics: a caller-owned device buffer the persistent kernels write phase timestamps into (slot 0: decode
// forward, slot 1: decode backward; 64 steps x 8 marks x 8 bytes each)
static std::atomic<unsigned long long*> g_debug{nullptr};
unsigned long long* debug_buffer(int slot) {
  unsigned long long* b = g_debug.load(std::memory_order_relaxed);
  return b ? b + (size_t)slot * 64 * 8 : nullptr;
}

}  // namespace l2s

extern "C" int l2s_set_debug_buffer(void* device_buffer, size_t bytes) {
  if (device_buffer != nullptr && bytes < 2 * 64 * 8 * sizeof(unsigned long long))
    return l2s::fail(L2S_ERR_ARG, "set_debug_buffer: need at least %zu bytes", 2 * 64 * 8 * sizeof(unsigned long long));
  l2s::g_debug.store(reinterpret_cast<unsigned long long*>(device_buffer), std::memory_order_relaxed);
  return L2S_OK;
}

extern "C" int l2s_version(void) { return 100; }
extern "C" const char* l2s_last_error_string(void) { return l2s::g_err; }
extern "C" uint64_t l2s_launch_count(void) { return l2s::g_launches.load(std::memory_order_relaxed); }

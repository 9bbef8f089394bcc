// Shared helpers for the lang2seg_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/l2s.h"

namespace l2s {

// ---- error plumbing (never exit(), never throw across the C ABI) -------------------------
void set_error(const char* fmt, ...);
int  fail(int code, const char* fmt, ...);

#define L2S_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::l2s::fail(L2S_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                              \
  } while (0)

#define L2S_LAUNCH_OK(name)                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::l2s::fail(L2S_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

#define L2S_REQUIRE(cond, code, ...)                                                       \
  do {                                                                                     \
    if (!(cond)) return ::l2s::fail(code, __VA_ARGS__);                                    \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int  sm_count();
int  max_smem_optin();
// Programmatic dependent launch for the chains of short kernels (decode loop, LSTM recurrence): the next kernel of
// the chain is scheduled while the current one runs and blocks in griddepcontrol.wait until that one has completed
// and flushed.  Opt-in (L2S_PDL=1): inside the whole-step CUDA graph, where kernel-to-kernel gaps are already
// sub-microsecond, it made the step SLOWER on B200 (11.05 vs 10.57 ms); it only pays for eager launches.
bool pdl_enabled();
bool env_flag(const char* name);   // getenv(name) starts with '1'; callers cache the result in a function-local static
void count_launch(int n = 1);          // feeds l2s_launch_count()
unsigned long long* debug_buffer(int slot);   // l2s_set_debug_buffer: phase timestamps of the persistent kernels, or null

// ---- internal launchers shared between translation units (att.cu <-> decode.cu) -----------
int launch_att_step_fwd(const float* att_h, int ldh, const float* att_feats, const float* p_att, const float* alpha_w,
                        const float* alpha_b, float* weight, float* att_res, int B, int A, int D, int Dh,
                        cudaStream_t st);
int launch_att_step_bwd(const float* datt_res, const float* att_h, int ldh, const float* att_feats, const float* p_att,
                        const float* alpha_w, const float* weight, float* datt_h, int ld_dah, float* de, float* dp_att,
                        float* datt_feats, float* dalpha_w, int B, int A, int D, int Dh, cudaStream_t st);
int launch_gates_fwd(const float* sums, int lds, const float* a2c_out, const float* c_prev, float* h, float* c, int B,
                     int D, cudaStream_t st);
int launch_gates_bwd(const float* sums, int lds, const float* a2c_out, const float* c_prev, const float* c,
                     const float* dh_a, const float* dh_b, const float* dc, float* dsums, int ld_ds, float* da2c,
                     float* dc_prev, int B, int D, cudaStream_t st);

// persistent decode kernels (decode_persist.cu): cluster size to launch with, 0 = shape / device not eligible
int decode_persist_cluster(int B, int A, int D, int Dh);
int launch_decode_fwd_persist(int cs, float* cat_all, const float* att, const float* p_att, const float* w_cat,
                              const float* w_a2c, const float* b_a2c, const float* alpha_w, const float* alpha_b,
                              float* h_all, float* c_all, float* a2c_all, float* pi_all, float* res_all, int T, int B,
                              int A, unsigned* bar, cudaStream_t st);

bool decode_bwd_persist_ok(int B, int A, int D, int Dh);
int launch_decode_bwd_persist(const float* dh_all, const float* cat_all, const float* att, const float* p_att,
                              const float* w_cat_t, const float* w_a2c_t, const float* alpha_w, const float* c_all,
                              const float* a2c_all, const float* pi_all, float* dcat_all, float* da2c_all,
                              float* dres_all, float* de_all, float* dh_carry, int T, int B, int A, unsigned* bar,
                              cudaStream_t st);

// tensor-core dynamic filter forward (dynfilter_tc.cu): 0 = ran, 1 = shape outside its range, < 0 = error
size_t dynfilter_tc_workspace_bytes(int E, int C, int HW);
int launch_dynfilter_tc_fwd(const float* X, const float* filt, const float* fuse, const int* e2i, float* response,
                            float* rk_saved, float* Y, const float* target, float* loss, int I, int E, int C, int H,
                            int W, int flags, void* workspace, size_t ws_bytes, cudaStream_t st);

// TMA-streamed dynamic filter backward (dynfilter_bwd_tma.cu): 0 = ran, 1 = shape outside its range, < 0 = error
size_t dynfilter_bwd_tma_workspace_bytes(int I);
int launch_dynfilter_bwd_tma(const float* X, const float* filt, const float* fuse, const int* e2i, const float* response,
                             const float* dY, const float* dresp, const float* target, const float* gscale, float* dX,
                             float* drbuf, int I, int E, int C, int H, int W, int flags, void* seg_ws, size_t seg_bytes,
                             cudaStream_t st);

// persistent bi-LSTM kernels (lstm_persist.cu)
bool bilstm_persist_ok(int L, int B, int H);
int launch_bilstm_fwd_persist(float* G, const float* w_hh, const int* lens, float* c_all, float* h_all, float* out,
                              float* hidden, int L, int B, unsigned* bar, cudaStream_t st);
int launch_bilstm_bwd_persist(const float* dout, const float* dhidden, const float* G, const float* w_hh_t,
                              const int* lens, const float* c_all, float* dG, int L, int B, unsigned* bar,
                              cudaStream_t st);

size_t linear_small_workspace_bytes(int M, int N, int K);
size_t linear_small_counter_bytes(int M, int N);
// D[M,N] (+)= A[M,K] . W[N,K]^T + bias ; workspace = [linear_small_counter_bytes(M,N) of zeroed counters | split-K partials]
int launch_linear_small(const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M,
                        int N, int K, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t st);

// ---- device helpers --------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// tanh with ~2 ulp error from ex2/rcp based evaluation: 1 - 2/(exp(2x)+1)
__device__ __forceinline__ float tanhf_fast_acc(float x) {
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch ----
// Every kernel launched through launch_chain() starts with pdl_launch_dependents() and executes pdl_wait() before its
// first access to memory that the previous kernel of the stream may write (and before its own first global write).
// Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// try_wait with a suspend-time hint (ns): the waiting warp is parked by the hardware until the phase completes or the
// hint expires, instead of burning issue slots in a tight retry loop next to the warps that do the work
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
  } while (!ok);
}

// ---- bulk async copies (TMA engine, 1-D) ----
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// pull a global range into L2 ahead of the shared-memory ring (decouples DRAM latency from the ring depth)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

template <class... KArgs, class... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__

}  // namespace l2s

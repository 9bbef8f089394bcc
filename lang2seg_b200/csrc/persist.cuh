// Shared pieces of the persistent recurrence kernels (decode_persist.cu, lstm_persist.cu): grid barrier on a global
// counter, register-blocked skinny-GEMM tile against shared-memory operands, cooperative + cluster launch helpers.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace l2s {
namespace {

constexpr int PD = 512;           // rnn_size == att_hid_size
constexpr int PG = 128;           // CTAs of the persistent grid
constexpr int PT = 256;           // threads per CTA
constexpr int PMAXB = 64;         // samples
constexpr int PQ = PD / 4;        // float4 per 512-float row
constexpr int PMAXLOC = 256;      // attention locations per CTA slice

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// Phase timestamps of CTA 0 (diagnostics: l2s_set_debug_buffer): slot k of step s -> prof[s * 8 + k], nanoseconds
struct PhaseProf {
  unsigned long long* buf;
  __device__ __forceinline__ void mark(int step, int k) const {
    if (buf != nullptr && blockIdx.x == 0 && threadIdx.x == 0) buf[step * 8 + k] = globaltimer_ns();
  }
};

// Grid barrier on one monotonically increasing counter (zeroed by the host before the launch).  arrive() and wait()
// are separate so that work which does not depend on the other CTAs can run in between.
struct GridBar {
  unsigned* ctr;
  unsigned epoch;
  unsigned n;          // CTAs that take part
  __device__ __forceinline__ void arrive() {
    __syncthreads();                     // every thread's global writes of this phase are ordered before the release
    if (threadIdx.x == 0) {
      __threadfence();
      red_release_gpu(ctr);
    }
    ++epoch;
  }
  __device__ __forceinline__ void wait() const {
    if (threadIdx.x == 0) {
      const unsigned target = epoch * n;
      while (ld_acquire_gpu(ctr) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
};

// transpose-reduce: lane l ends with the sum over all lanes of v[l]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float a = v[i], b = v[i + s];
      const float keep = up ? b : a, send = up ? a : b;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ float dot4(const float4& a, const float4& w, float acc) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  return fmaf(a.w, w.w, acc);
}

// One register-blocked tile of a skinny GEMM: out[r*NC + c] = sum_k A[row_r][k] * W[col_c][k] for R rows and NC columns
// held in shared memory (row stride lda4 / ldw4 float4), K = 128 * KS (lane l owns float4 l + 32 ks of every row).
// On return lane (r*NC + c) holds the total of output (r, c).
template <int R, int NC, int KS>
__device__ __forceinline__ float gemv_tile(const float4* __restrict__ sA4, int lda4, const int (&arow)[R],
                                           const float4* __restrict__ sW4, int ldw4, int wrow0, int lane) {
  static_assert(R * NC <= 32, "tile too large for one transpose-reduce");
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    float4 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = sA4[arow[r] * lda4 + ks * 32 + lane];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 w = sW4[(wrow0 + c) * ldw4 + ks * 32 + lane];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r * NC + c] = dot4(a[r], w, acc[r * NC + c]);
    }
  }
  return transpose_reduce32(acc, lane);
}

// rows [0,B) x 512 floats from global (produced by other CTAs of this launch: L2 loads, never L1) into shared memory
__device__ __forceinline__ void stage_rows(float4* __restrict__ sA4, const float* __restrict__ src, int B, int ld) {
  for (int i = threadIdx.x; i < B * PQ; i += PT) {
    const int b = i / PQ, q = i - b * PQ;
    sA4[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)b * ld) + q);
  }
}

// rows [0,B) x (4 nq) floats starting at src (row stride ld) -> shared memory [B][nq] float4
__device__ __forceinline__ void stage_slice(float4* __restrict__ sA4, const float* __restrict__ src, int B, int ld, int nq) {
  for (int i = threadIdx.x; i < B * nq; i += PT) {
    const int b = i / nq, q = i - b * nq;
    sA4[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)b * ld) + q);
  }
}

template <class Kern, class Args>
int launch_persistent(Kern kern, int cs, size_t smem, cudaStream_t st, const Args& args, bool coop) {
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PG);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop ? 2 : 1;
  L2S_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args));
  count_launch();
  return L2S_OK;
}

// how many clusters of `cs` CTAs of this kernel can be resident at once (0 on error)
template <class Kern>
int max_clusters(Kern kern, int cs, size_t smem) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PG);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace
}  // namespace l2s

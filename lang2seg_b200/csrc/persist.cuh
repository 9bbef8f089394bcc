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
constexpr int PT = 512;           // threads per CTA (16 warps: four per scheduler hide the LDS / FFMA latencies of the tiles)
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
    __syncthreads();                     // every thread's global writes of this phase happen before thread 0's release
    if (threadIdx.x == 0) red_release_gpu(ctr);      // release at gpu scope is cumulative over the barrier above
    ++epoch;
  }
  __device__ __forceinline__ void wait() const {
    if (threadIdx.x == 0) {
      const unsigned target = epoch * n;
      while (ld_acquire_gpu(ctr) < target) {
      }
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

// Operand staging global -> shared through the TMA engine (1-D bulk copies completing on an mbarrier): one elected
// thread (or warp, for strided rows) issues, everybody waits on the barrier.  Measured against a plain __ldcg loop
// (12 loads per thread for a 96 KB block = three exposed L2 round trips): 3.0 vs 4.5 us for stage + first tile at
// B = 48 (profiles/r02e_decode_phases_*).  The data was written by OTHER CTAs of this launch and acquired through the
// grid barrier; fence.proxy.async orders those generic-proxy observations before the async-proxy reads.  Callers
// guarantee (CTA barrier) that nobody still reads the destination.
struct Stager {
  uint64_t* bar;
  unsigned phase;
  __device__ __forceinline__ void init() {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    phase = 0;
  }
  // [bytes] contiguous (multiple of 16)
  __device__ __forceinline__ void load_contig(void* sdst, const void* gsrc, uint32_t bytes) {
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_arrive_expect_tx(bar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u)
        bulk_g2s(reinterpret_cast<char*>(sdst) + off, reinterpret_cast<const char*>(gsrc) + off,
                 bytes - off < 32768u ? bytes - off : 32768u, bar);
    }
  }
  // B rows of row_bytes each (multiple of 16), global row stride gstride bytes, packed in shared memory
  __device__ __forceinline__ void load_rows(void* sdst, const void* gsrc, int B, uint32_t row_bytes, size_t gstride) {
    if (threadIdx.x < 32) {
      asm volatile("fence.proxy.async;" ::: "memory");
      if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)B * row_bytes);
      __syncwarp();
      for (int b = threadIdx.x; b < B; b += 32)
        bulk_g2s(reinterpret_cast<char*>(sdst) + (size_t)b * row_bytes, reinterpret_cast<const char*>(gsrc) + (size_t)b * gstride,
                 row_bytes, bar);
    }
  }
  __device__ __forceinline__ void wait() {
    mbar_wait(bar, phase & 1u);
    phase ^= 1u;
  }
};

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

// Probe: what does ONE SM sustain with 1-D bulk copies (cp.async.bulk, SASS UBLKCP) vs cp.async (LDGSTS)?
// Each CTA (1 per SM) streams its own slice of a large buffer through a shared-memory ring; nothing touches the data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bulk_probe tma_bulk_probe.cu && ./tma_bulk_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

constexpr int kMaxStages = 64;

// mode 0: one bulk copy of `chunk` bytes per stage.  mode 1: `split` lanes each copy chunk/split bytes.
__global__ void __launch_bounds__(64, 1) load_bulk(const char* __restrict__ src, size_t per_cta, int chunk, int stages, int split) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  const int t = threadIdx.x;
  if (t == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  if (t < 32) {
    const int part = chunk / split;
    for (int k = 0; k < n; ++k) {
      const int s = k % stages;
      if (t == 0) {
        if (k >= stages) mbar_wait(&empty[s], ((k / stages) - 1) & 1);
        mbar_expect(&full[s], (uint32_t)chunk);
      }
      __syncwarp();
      if (t < split) bulk_g2s(smem + (size_t)s * chunk + t * part, base + (size_t)k * chunk + t * part, (uint32_t)part, &full[s]);
    }
  } else if (t == 32) {
    for (int k = 0; k < n; ++k) {
      const int s = k % stages;
      mbar_wait(&full[s], (k / stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

// cp.async (LDGSTS) producer warp: 16 bytes per lane per instruction, completion through the same mbarrier
__global__ void __launch_bounds__(64, 1) load_ldgsts(const char* __restrict__ src, size_t per_cta, int chunk, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  const int t = threadIdx.x;
  if (t == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 32); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  if (t < 32) {
    for (int k = 0; k < n; ++k) {
      const int s = k % stages;
      if (k >= stages) mbar_wait(&empty[s], ((k / stages) - 1) & 1);
      for (int o = t * 16; o < chunk; o += 512)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + (size_t)s * chunk + o)), "l"(base + (size_t)k * chunk + o) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
  } else if (t == 32) {
    for (int k = 0; k < n; ++k) {
      const int s = k % stages;
      mbar_wait(&full[s], (k / stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

// bulk stores smem -> global, `depth` groups in flight (wait_group.read keeps depth-1 pending)
template <int DEPTH>
__global__ void __launch_bounds__(64, 1) store_bulk(char* __restrict__ dst, size_t per_cta, int chunk, int per_group) {
  extern __shared__ __align__(128) unsigned char smem[];
  char* base = dst + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  if (threadIdx.x == 0) {
    for (int k = 0; k < n; k += per_group) {
      for (int j = 0; j < per_group && k + j < n; ++j)
        bulk_s2g(base + (size_t)(k + j) * chunk, smem + (size_t)(((k / per_group) % DEPTH) * per_group + j) * chunk, (uint32_t)chunk);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// plain coalesced stores from registers (what a non-TMA epilogue would do): 4 warps, float4 per lane
__global__ void __launch_bounds__(1024, 1) store_st(char* __restrict__ dst, size_t per_cta) {
  float4* base = reinterpret_cast<float4*>(dst + (size_t)blockIdx.x * per_cta);
  const size_t n = per_cta / 16;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) __stcs(base + i, v);
}

template <class F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = (size_t)6272 * 3072;      // 19.3 MB per CTA, 2.85 GB in total: larger than L2
  const size_t total = per_cta * sms;
  char* buf;
  if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, total);
  const int maxsm = 200 * 1024;
  cudaFuncSetAttribute(load_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
  cudaFuncSetAttribute(load_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
  cudaFuncSetAttribute(store_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
  cudaFuncSetAttribute(store_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
  printf("SMs %d, %.2f GB per pass\n", sms, total / 1e9);
  struct V { int chunk, stages, split; };
  const V vs[] = {{6272, 8, 1}, {6272, 12, 1}, {6272, 24, 1}, {6272, 12, 7}, {6272, 24, 7}, {12544, 12, 1}, {25088, 6, 1}, {1568, 48, 1}, {896, 64, 1}};
  for (const V& v : vs) {
    const size_t smem = (size_t)v.chunk * v.stages;
    float ms = time_ms([&] { load_bulk<<<sms, 64, smem>>>(buf, per_cta, v.chunk, v.stages, v.split); });
    printf("load  bulk   chunk %6d stages %2d split %d : %7.3f ms  %7.1f GB/s  %5.2f B/clk/SM@1.9GHz  err=%s\n", v.chunk, v.stages, v.split, ms,
           total / ms / 1e6, total / (double)sms / (ms * 1e-3 * 1.9e9), cudaGetErrorString(cudaGetLastError()));
  }
  for (const V& v : vs) {
    if (v.split != 1 || v.chunk % 512) { if (v.chunk != 6272) continue; }
    if (v.split != 1) continue;
    const size_t smem = (size_t)v.chunk * v.stages;
    float ms = time_ms([&] { load_ldgsts<<<sms, 64, smem>>>(buf, per_cta, v.chunk, v.stages); });
    printf("load  ldgsts chunk %6d stages %2d         : %7.3f ms  %7.1f GB/s  %5.2f B/clk/SM@1.9GHz  err=%s\n", v.chunk, v.stages, ms, total / ms / 1e6,
           total / (double)sms / (ms * 1e-3 * 1.9e9), cudaGetErrorString(cudaGetLastError()));
  }
  for (int pg : {1, 4, 8}) {
    float ms = time_ms([&] { store_bulk<2><<<sms, 64, (size_t)2 * pg * 6272>>>(buf, per_cta, 6272, pg); });
    printf("store bulk   chunk   6272 depth 2 x %d/group : %7.3f ms  %7.1f GB/s  err=%s\n", pg, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    ms = time_ms([&] { store_bulk<4><<<sms, 64, (size_t)4 * pg * 6272>>>(buf, per_cta, 6272, pg); });
    printf("store bulk   chunk   6272 depth 4 x %d/group : %7.3f ms  %7.1f GB/s  err=%s\n", pg, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  {
    float ms = time_ms([&] { store_st<<<sms, 1024>>>(buf, per_cta); });
    printf("store st.cs  float4, 1024 threads            : %7.3f ms  %7.1f GB/s\n", ms, total / ms / 1e6);
  }
  return 0;
}

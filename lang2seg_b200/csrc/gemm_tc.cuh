// fp32-accurate GEMM on the Blackwell bf16 tensor pipe (tcgen05 + TMEM + TMA), sm_100a.
//
// D[M,N] = A[M,K] * B[N,K]^T with every fp32 operand pre-split into bf16 planes (hi, lo):
//     acc += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (fp32 accumulation in TMEM)
// The dropped terms are O(2^-16) relative, so results sit ~1e-5 from an fp32 GEMM -- inside the
// 1e-4 parity budget that rules out single-pass TF32/bf16 (hard parts).
//
// Persistent, warp-specialised CTA (192 threads):
//   warp 0      TMA producer: 4-stage ring of {A_hi, A_lo, B_hi, B_lo} tiles, one mbarrier per stage
//   warp 1      allocates TMEM, single elected lane issues tcgen05.mma (kind::f16, M=128, N=BN),
//               tcgen05.commit releases smem stages and publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator (two TMEM buffers, so the epilogue of
//               tile i overlaps the main loop of tile i+1) and apply an Epi functor
// Operands may be K-major ([rows][K], 64B-swizzled TMA boxes) or MN-major ([K][rows], 128B
// swizzle) so the weight-gradient GEMMs read activations in place.  M/N/K tails rely on TMA
// out-of-bounds zero fill; the epilogue guards rows/cols.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace l2s {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;        // bf16 elements per k-block (64 bytes)
constexpr int STAGES = 4;
constexpr int THREADS = 192;

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* t) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(t) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (sm_100 UMMA): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t kLayoutSW128 = 2, kLayoutSW64 = 4;

__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  // c_format F32 (1) @4, a/b format BF16 (1) @7/@10, majors @15/@16, N>>3 @17, M>>4 @24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN>
struct SmemPlan {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t TOTAL = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ Maps maps, int M, int N, int K, int nsplit, int kb_per_split, Epi epi) {
  using P = SmemPlan<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * P::STAGE);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int kblocks = (K + BK - 1) / BK;
  const int total = tiles_m * tiles_n * nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);     // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      prefetch_tmap(&maps.a_hi); prefetch_tmap(&maps.a_lo); prefetch_tmap(&maps.b_hi); prefetch_tmap(&maps.b_lo);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int split = tile % nsplit;
        const int tn = (tile / nsplit) % tiles_n;
        const int tm = tile / (nsplit * tiles_n);
        const int kb0 = split * kb_per_split, kb1 = min(kblocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
          uint8_t* st = tiles + s * P::STAGE;
          mbar_arrive_expect_tx(&full[s], P::STAGE);
          const int k0 = kb * BK, m0 = tm * BM, n0 = tn * BN;
          if (!A_MN) {
            tma_load_2d(st, &maps.a_hi, k0, m0, &full[s]);
            tma_load_2d(st + P::A_BYTES, &maps.a_lo, k0, m0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              tma_load_2d(st + j * (BK * 128), &maps.a_hi, m0 + 64 * j, k0, &full[s]);
              tma_load_2d(st + P::A_BYTES + j * (BK * 128), &maps.a_lo, m0 + 64 * j, k0, &full[s]);
            }
          }
          uint8_t* sb = st + 2 * P::A_BYTES;
          if (!B_MN) {
            tma_load_2d(sb, &maps.b_hi, k0, n0, &full[s]);
            tma_load_2d(sb + P::B_BYTES, &maps.b_lo, k0, n0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              tma_load_2d(sb + j * (BK * 128), &maps.b_hi, n0 + 64 * j, k0, &full[s]);
              tma_load_2d(sb + P::B_BYTES + j * (BK * 128), &maps.b_lo, n0 + 64 * j, k0, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int split = tile % nsplit;
      const int kb0 = split * kb_per_split, kb1 = min(kblocks, kb0 + kb_per_split);
      const int acc = tcount & 1;
      if (tcount >= 2) mbar_wait(&tempty[acc], ((tcount >> 1) - 1) & 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = base + s * P::STAGE, sb = sa + 2 * P::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            uint64_t ahi, alo, bhi, blo;
            if (!A_MN) {
              ahi = make_desc(sa + kk * 32, 0, 512, kLayoutSW64);
              alo = make_desc(sa + P::A_BYTES + kk * 32, 0, 512, kLayoutSW64);
            } else {
              ahi = make_desc(sa + kk * 2048, BK * 128, 1024, kLayoutSW128);
              alo = make_desc(sa + P::A_BYTES + kk * 2048, BK * 128, 1024, kLayoutSW128);
            }
            if (!B_MN) {
              bhi = make_desc(sb + kk * 32, 0, 512, kLayoutSW64);
              blo = make_desc(sb + P::B_BYTES + kk * 32, 0, 512, kLayoutSW64);
            } else {
              bhi = make_desc(sb + kk * 2048, BK * 128, 1024, kLayoutSW128);
              blo = make_desc(sb + P::B_BYTES + kk * 2048, BK * 128, 1024, kLayoutSW128);
            }
            umma_f16(d_tmem, alo, bhi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            umma_f16(d_tmem, ahi, blo, idesc, 1u);
            umma_f16(d_tmem, ahi, bhi, idesc, 1u);
          }
          umma_commit(&empty[s]);                 // smem stage free once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =================
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int split = tile % nsplit;
      const int tn = (tile / nsplit) % tiles_n;
      const int tm = tile / (nsplit * tiles_n);
      const int acc = tcount & 1;
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      fence_after_sync();
      const int row = tm * BM + q * 32 + lane;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = tn * BN + ch * 32;
        if (col0 >= N) break;             // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + ch * 32, v);
        epi(row, col0, v, M, N, split);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

// bf16 operand plane.  K-major: memory [rows][K] (ld elements per row).  MN-major: memory [K][rows].
int make_operand_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K, int64_t ld, bool mn_major,
                     int box_rows);

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                int64_t ldb, int M, int N, int K, int split_k, const Epi& epi, cudaStream_t st) {
  Maps maps;
  int rc;
  if ((rc = make_operand_map(&maps.a_hi, a_hi, M, K, lda, A_MN, BM))) return rc;
  if ((rc = make_operand_map(&maps.a_lo, a_lo, M, K, lda, A_MN, BM))) return rc;
  if ((rc = make_operand_map(&maps.b_hi, b_hi, N, K, ldb, B_MN, BN))) return rc;
  if ((rc = make_operand_map(&maps.b_lo, b_lo, N, K, ldb, B_MN, BN))) return rc;
  const int kblocks = (K + BK - 1) / BK;
  split_k = split_k < 1 ? 1 : (split_k > kblocks ? kblocks : split_k);
  const int kb_per = (kblocks + split_k - 1) / split_k;
  const int nsplit = (kblocks + kb_per - 1) / kb_per;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * nsplit;
  auto kern = gemm_bf16x3_kernel<BN, A_MN, B_MN, Epi>;
  const size_t smem = SmemPlan<BN>::TOTAL;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, THREADS, smem, st>>>(maps, M, N, K, nsplit, kb_per, epi);
  L2S_LAUNCH_OK("gemm_bf16x3_kernel");
  count_launch();
  return L2S_OK;
}

}  // namespace tc
}  // namespace l2s

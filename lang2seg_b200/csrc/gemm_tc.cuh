// fp32-accurate GEMM on the Blackwell bf16 tensor pipe (tcgen05 + TMEM + TMA), sm_100a.
//
// D[M,N] = A[M,K] * B[N,K]^T with every fp32 operand pre-split into bf16 planes (hi, lo):
//     acc += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (fp32 accumulation in TMEM)
// The dropped terms are O(2^-16) relative, so results sit ~1e-5 from an fp32 GEMM -- inside the
// 1e-4 parity budget that rules out single-pass TF32/bf16 (hard parts).
// With l2s_set_precision(L2S_PRECISION_BF16) the same kernel issues the A_hi*B_hi pass only and its producer leaves
// the lo planes alone: the bf16 variant (1e-2 tolerance).
//
// Persistent, warp-specialised CTA (64 + 128*MH*EW threads), CTA tile = (128*MH) x BN:
//   warp 0      TMA producer: ring of {A_hi, A_lo, B_hi, B_lo} tiles, one mbarrier per stage
//   warp 1      allocates TMEM, single elected lane issues tcgen05.mma (kind::f16, M=128, N=BN),
//               tcgen05.commit releases smem stages and publishes the accumulator
//   warps 2..   epilogue (4 warps per 128-row half): tcgen05.ld the fp32 accumulator (lane = row, 32 columns per
//               load), apply an Epi functor.  Row-major outputs go through a per-warp 32x33 shared-memory transpose
//               tile so that every global store instruction writes whole 128-byte lines (a lane-per-row store
//               touches 32 sectors per instruction).
// MH = 1: 128 x BN tile, 4 stages, two TMEM accumulators (the epilogue of tile i overlaps the main loop of
//         tile i+1).  Operand traffic: 24 KB of L2->SM bytes per 3 MMAs (BN = 256) = 64 B/clk/SM at the
//         tensor peak, above the ~42 B/clk/SM the L2 can deliver chip-wide -- this shape is L2-bound.
// MH = 2: 256 x BN tile as two 128-row accumulators that share every B tile: 32 KB per 6 MMAs = 42.7 B/clk/SM.
//         3 stages of 64 KB; the two accumulators fill TMEM (2 x 256 columns), so the epilogue is not
//         overlapped -- 8 epilogue warps keep it short (a few % of a K = 2048 main loop).
// EW = 2 (with MH = 1): two epilogue warps per TMEM lane quadrant, splitting the column chunks -- for the GEMMs
//         whose K is so short that the epilogue (global loads / scattered stores) is the critical path.
// Operands may be K-major ([rows][K], 64B-swizzled TMA boxes) or MN-major ([K][rows], 128B
// swizzle) so the weight-gradient GEMMs read activations in place.  M/N/K tails rely on TMA
// out-of-bounds zero fill; the epilogue guards rows/cols.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace l2s {
namespace tc {

constexpr int BM = 128;       // rows per MMA / per accumulator
constexpr int BK = 32;        // bf16 elements per k-block (64 bytes)

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

template <int BN, int MH, int EW = 1>
struct SmemPlan {
  static constexpr int STAGES = MH == 1 ? 4 : 3;
  static constexpr int EPI_WARPS = 4 * MH * EW;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int NACC = MH == 1 ? 2 : 1;       // TMEM accumulator buffers (each MH * BN columns)
  static constexpr uint32_t HALF_BYTES = BM * BK * 2;          // one 128-row half of an A plane
  static constexpr uint32_t A_BYTES = MH * HALF_BYTES;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t SCRATCH = 32 * 33 * 4;             // per epilogue warp: a 32 x 32 fp32 transpose tile (padded)
  static constexpr uint32_t TOTAL = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_WARPS * SCRATCH;
};

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

// Work item -> (tm, tn, split).  raster 0: the K splits of one output tile are adjacent (the co-resident CTAs then work on
// ~16 output tiles x all splits).  raster 1 (split-K launches): split-major -- the co-resident CTAs work on ALL output
// tiles of a few K ranges, so every A panel is shared by tiles_n CTAs and every B panel by tiles_m CTAs while it is hot
// in L2 (the weight-gradient GEMM read its operands 1.96x from DRAM with raster 0, profiles/r01).
__device__ __forceinline__ void decode_tile(int tile, int tiles_m, int tiles_n, int nsplit, int raster, int* tm, int* tn,
                                            int* split) {
  if (raster == 0) {
    *split = tile % nsplit;
    *tn = (tile / nsplit) % tiles_n;
    *tm = tile / (nsplit * tiles_n);
  } else {
    const int mn = tiles_m * tiles_n;
    *split = tile / mn;
    const int r = tile - *split * mn;
    *tn = r % tiles_n;
    *tm = r / tiles_n;
  }
}

template <int BN, int MH, int EW, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(64 + 128 * MH * EW, 1)
gemm_bf16x3_kernel(const __grid_constant__ Maps maps, int M, int N, int K, int nsplit, int kb_per_split, int passes,
                   int raster, Epi epi) {
  using P = SmemPlan<BN, MH, EW>;
  constexpr int STAGES = P::STAGES;
  constexpr int NACC = P::NACC;
  constexpr int TM = BM * MH;            // rows per CTA tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * P::STAGE);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2] (NACC used)
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* scratch_all = reinterpret_cast<float*>(tiles + STAGES * P::STAGE + 256);   // [EPI_WARPS][32][33]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + TM - 1) / TM, tiles_n = (N + BN - 1) / BN;
  const int kblocks = (K + BK - 1) / BK;
  const int total = tiles_m * tiles_n * nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], P::EPI_WARPS);     // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NACC * MH * BN);
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
        int split, tn, tm;
        decode_tile(tile, tiles_m, tiles_n, nsplit, raster, &tm, &tn, &split);
        const int kb0 = split * kb_per_split, kb1 = min(kblocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
          uint8_t* st = tiles + s * P::STAGE;
          const bool lo = passes == 3;        // the single-pass bf16 variant never touches the lo planes
          mbar_arrive_expect_tx(&full[s], lo ? P::STAGE : P::STAGE / 2);
          const int k0 = kb * BK, m0 = tm * TM, n0 = tn * BN;
          if (!A_MN) {
            tma_load_2d(st, &maps.a_hi, k0, m0, &full[s]);
            if (lo) tma_load_2d(st + P::A_BYTES, &maps.a_lo, k0, m0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < TM / 64; ++j) {
              tma_load_2d(st + j * (BK * 128), &maps.a_hi, m0 + 64 * j, k0, &full[s]);
              if (lo) tma_load_2d(st + P::A_BYTES + j * (BK * 128), &maps.a_lo, m0 + 64 * j, k0, &full[s]);
            }
          }
          uint8_t* sb = st + 2 * P::A_BYTES;
          if (!B_MN) {
            tma_load_2d(sb, &maps.b_hi, k0, n0, &full[s]);
            if (lo) tma_load_2d(sb + P::B_BYTES, &maps.b_lo, k0, n0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              tma_load_2d(sb + j * (BK * 128), &maps.b_hi, n0 + 64 * j, k0, &full[s]);
              if (lo) tma_load_2d(sb + P::B_BYTES + j * (BK * 128), &maps.b_lo, n0 + 64 * j, k0, &full[s]);
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
      int split, tn_unused, tm_unused;
      decode_tile(tile, tiles_m, tiles_n, nsplit, raster, &tm_unused, &tn_unused, &split);
      const int kb0 = split * kb_per_split, kb1 = min(kblocks, kb0 + kb_per_split);
      const int acc = tcount % NACC;
      if (tcount >= NACC) mbar_wait(&tempty[acc], ((tcount / NACC) - 1) & 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * MH * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = base + s * P::STAGE, sb = sa + 2 * P::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            uint64_t bhi, blo;
            if (!B_MN) {
              bhi = make_desc(sb + kk * 32, 0, 512, kLayoutSW64);
              blo = make_desc(sb + P::B_BYTES + kk * 32, 0, 512, kLayoutSW64);
            } else {
              bhi = make_desc(sb + kk * 2048, BK * 128, 1024, kLayoutSW128);
              blo = make_desc(sb + P::B_BYTES + kk * 2048, BK * 128, 1024, kLayoutSW128);
            }
#pragma unroll
            for (int h = 0; h < MH; ++h) {
              // a 128-row half of an A plane is HALF_BYTES (8 KB) in both layouts
              const uint32_t sh = sa + h * P::HALF_BYTES;
              uint64_t ahi, alo;
              if (!A_MN) {
                ahi = make_desc(sh + kk * 32, 0, 512, kLayoutSW64);
                alo = make_desc(sh + P::A_BYTES + kk * 32, 0, 512, kLayoutSW64);
              } else {
                ahi = make_desc(sh + kk * 2048, BK * 128, 1024, kLayoutSW128);
                alo = make_desc(sh + P::A_BYTES + kk * 2048, BK * 128, 1024, kLayoutSW128);
              }
              const uint32_t d_half = d_tmem + h * BN;
              const uint32_t acc0 = (kb > kb0 || kk > 0) ? 1u : 0u;
              if (passes == 3) {
                umma_f16(d_half, alo, bhi, idesc, acc0);
                umma_f16(d_half, ahi, blo, idesc, 1u);
                umma_f16(d_half, ahi, bhi, idesc, 1u);
              } else {
                umma_f16(d_half, ahi, bhi, idesc, acc0);
              }
            }
          }
          umma_commit(&empty[s]);                 // smem stage free once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (warps 2..5 [, 6..9] -> TMEM lane quadrants 2,3,0,1 of half 0 [, 1]) =================
    // warp set = (warp - 2) / 4: the 128-row half it serves (MH = 2) or its share of the column chunks (EW = 2)
    const int q = warp & 3, wset = (warp - 2) >> 2;
    const int h = wset % MH, cs = wset / MH;
    float* scratch = scratch_all + (size_t)(warp - 2) * (P::SCRATCH / 4);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      int split, tn, tm;
      decode_tile(tile, tiles_m, tiles_n, nsplit, raster, &tm, &tn, &split);
      const int acc = tcount % NACC;
      mbar_wait(&tfull[acc], (tcount / NACC) & 1);
      fence_after_sync();
      const int row = tm * TM + h * BM + q * 32 + lane;
#pragma unroll 1
      for (int ch = cs; ch < BN / 32; ch += EW) {
        const int col0 = tn * BN + ch * 32;
        if (col0 >= N) break;             // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MH + h) * BN + ch * 32, v);
        epi(row, col0, v, M, N, split, scratch);
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
    tmem_dealloc(tmem_base, NACC * MH * BN);
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

// tensor passes per product: 3 = bf16x3 split (fp32 accurate, the default), 1 = hi planes only (the bf16 variant of
// north_star's tolerance clause: 1e-2).  Process-wide switch, l2s_set_precision() in include/l2s.h.
int gemm_passes();
// split-K work-item order: true = split-major (default), L2S_GEMM_RASTER=0 restores tile-major (A/B diagnostics)
bool split_raster();

// bf16 operand plane.  K-major: memory [rows][K] (ld elements per row).  MN-major: memory [K][rows].
int make_operand_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K, int64_t ld, bool mn_major,
                     int box_rows);

// split-K factor that fills the machine: minimises  waves * (k-blocks per split + epilogue) ; `epi_kb` is the cost of
// one (atomic) epilogue in k-block units
int auto_split(int tiles, int kblocks, int epi_kb);

template <int BN, int MH, int EW, bool A_MN, bool B_MN, class Epi>
int launch_gemm_mh(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                   int64_t ldb, int M, int N, int K, int split_k, const Epi& epi, cudaStream_t st) {
  using P = SmemPlan<BN, MH, EW>;
  Maps maps;
  int rc;
  // K-major A: one box of all 128*MH rows; MN-major boxes are always 64 rows wide
  if ((rc = make_operand_map(&maps.a_hi, a_hi, M, K, lda, A_MN, BM * MH))) return rc;
  if ((rc = make_operand_map(&maps.a_lo, a_lo, M, K, lda, A_MN, BM * MH))) return rc;
  if ((rc = make_operand_map(&maps.b_hi, b_hi, N, K, ldb, B_MN, BN))) return rc;
  if ((rc = make_operand_map(&maps.b_lo, b_lo, N, K, ldb, B_MN, BN))) return rc;
  const int kblocks = (K + BK - 1) / BK;
  const int tiles_mn = ((M + BM * MH - 1) / (BM * MH)) * ((N + BN - 1) / BN);
  if (split_k == 0) split_k = auto_split(tiles_mn, kblocks, 8 * MH);
  split_k = split_k < 1 ? 1 : (split_k > kblocks ? kblocks : split_k);
  const int kb_per = (kblocks + split_k - 1) / split_k;
  const int nsplit = (kblocks + kb_per - 1) / kb_per;
  const int tiles = tiles_mn * nsplit;
  auto kern = gemm_bf16x3_kernel<BN, MH, EW, A_MN, B_MN, Epi>;
  const size_t smem = P::TOTAL;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, P::THREADS, smem, st>>>(maps, M, N, K, nsplit, kb_per, gemm_passes(), (nsplit > 1 && split_raster()) ? 1 : 0, epi);
  L2S_LAUNCH_OK("gemm_bf16x3_kernel");
  count_launch();
  return L2S_OK;
}

// CTA shape of a launch: 0 = (128 rows, 4 epilogue warps), 1 = (128 rows, 8 epilogue warps), 2 = (256 rows, 8 warps)
enum Shape { kShapeAuto = -1, kShape128 = 0, kShape128E2 = 1, kShape256 = 2 };
// the L2S_GEMM_SHAPE environment override (diagnostics / parity tests), -1 = none
int forced_shape();

// split_k: >= 1 as given (the epilogue must then accumulate atomically), 0 = choose (split-K epilogues only).
// shape: kShapeAuto picks by the measured behaviour on B200 (profiles/): 256-row tiles pay off when a work item's
// K range is long (split-K weight gradients: the un-overlapped epilogue is amortised and a 256 x 256 tile moves 2/3
// of the L2->SM bytes per flop); 8 epilogue warps when K is so short that the epilogue is the critical path.
template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                int64_t ldb, int M, int N, int K, int split_k, const Epi& epi, cudaStream_t st, int shape = kShapeAuto) {
  const int kblocks = (K + BK - 1) / BK;
  if (shape == kShapeAuto) {
    shape = kShape128E2;        // 8 epilogue warps: with 4 the epilogue of a 128 x 256 tile outlasts a K <= 2048 main loop
    if (BN == 256 && split_k == 0 && M >= 2 * BM) {
      const int tn = (N + BN - 1) / BN;
      const int t2 = ((M + 2 * BM - 1) / (2 * BM)) * tn;
      const int s2 = auto_split(t2, kblocks, 16);
      if (kblocks / s2 >= 128) shape = kShape256;
    }
  }
  if (forced_shape() >= 0) shape = forced_shape();
  if (BN != 256 && shape == kShape256) shape = kShape128;
  if (shape == kShape256)
    return launch_gemm_mh<BN, (BN == 256 ? 2 : 1), 1, A_MN, B_MN, Epi>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, split_k, epi, st);
  if (shape == kShape128E2)
    return launch_gemm_mh<BN, 1, 2, A_MN, B_MN, Epi>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, split_k, epi, st);
  return launch_gemm_mh<BN, 1, 1, A_MN, B_MN, Epi>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, split_k, epi, st);
}

}  // namespace tc
}  // namespace l2s

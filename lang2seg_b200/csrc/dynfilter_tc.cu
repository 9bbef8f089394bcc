// Spatial dynamic-filter response layer, forward, on the Blackwell tensor pipe (north_star kernel (1)).
//
// Semantics: network_cycle_response.py:534-570 + response loss :415-422, identical to
// dynfilter.cu:  r_k[p] = M_k[p] sum_c f_k[c] X[c,p] ;  r = sum_k w_k r_k ;  Y = X * sigmoid(r).
//
// CTA = (image, tile of TPX pixels), 8 warps:
//   * the [C x TPX px] fp32 tile of X is brought in ONCE by TMA (2-D tensor map over (HW, I*C), boxes of 256
//     channels; TPX = 32: 128-byte rows = whole DRAM lines) and stays in shared memory: the contraction AND the gating
//     read it from there, so X costs its algorithmic bytes and nothing else;
//   * the contraction  D[(e,k), p] = sum_c f_{e,k}[c] X[c,p]  for the expressions of the image runs on tcgen05.mma
//     (kind::f16, fp32 accumulators in TMEM) with both operands split into bf16 (hi, lo) planes (~2^-16 relative,
//     inside the 1e-4 fp32 contract that rules out one-pass TF32/bf16).  The planes are STACKED inside the operand
//     tiles -- A = [f_hi (24 rows); f_lo (24 rows)], B = [X_hi (TPX rows); X_lo (TPX rows)] -- so ONE
//     M128 x N(2 TPX) x K16 instruction per 16 channels yields all four partial products (hi*hi, hi*lo, lo*hi, lo*lo in
//     the four (row block, column block) corners of the accumulator; the drain adds them up): a third of the
//     instructions of the three-pass scheme, which matters because an MMA this narrow is paced by its shared-memory
//     operand reads, not by the tensor math.  A stage of the ring is 64 channels: the filter planes arrive by TMA
//     (128B-swizzled K-major boxes of 24 rows); the X operand is produced from the resident fp32 tile by two groups of
//     four converter warps (alternating stages), which write the planes straight into the swizzled UMMA layout
//     (16-byte chunks, packed cvt.rn.bf16x2) and hand the stage to the MMA warp through ONE mbarrier per stage (TMA
//     transaction bytes + one arrive per converter warp);
//   * epilogue: two warps drain the accumulator rows from TMEM into shared memory, then thread = (expression, pixel)
//     applies the partition masks (comparisons on (y, x), no table), writes r_k (kept for the backward), the fused
//     response, the BCE-with-logits partial per (expression, tile) (summed in fixed order by a second tiny kernel:
//     deterministic, no float atomics) and the gate;
//   * gating: Y_e = X * gate_e streamed out with float4 st.cs for every expression of the image.
// TPX = 16 halves the tile (64 KB): two CTAs per SM, so that one CTA's load / contraction overlaps the other's stores.
// The filter block of a chunk is a 24-row TMA box (3 expressions x 7 filters), so an image with more than 3
// expressions is processed in chunks of 3 against the same resident tile (the reference's batches carry 1-3
// expressions per image).  Requirements (otherwise dynfilter.cu's FFMA kernel runs): H*W % 4 == 0 (TMA row stride),
// C % 64 == 0, C <= 1024 (tile residency), C <= 256 or C % 256 == 0 (whole TMA boxes of X).
#include <cuda.h>
#include <cuda_bf16.h>

#include "gemm_tc.cuh"

namespace l2s {
namespace {

constexpr int NF = L2S_NUM_FILTERS;
constexpr int DBK = 64;                 // channels per stage (128-byte bf16 rows, SW128)
constexpr int DTHREADS = 320;           // producer warp, MMA warp, 2 x 4 converter warps
constexpr int ABOX = 24;                // filter rows per TMA box
constexpr int DMAXE = ABOX / NF;        // expressions per chunk: 7 * 3 = 21 <= 24 box rows
constexpr int XBOX = 256;               // channels per TMA box of X
constexpr uint32_t A_PLANE = ABOX * 128;   // bytes per A plane of a stage actually filled by TMA (24 rows x 128 B = 3 swizzle
                                           // atoms); the descriptor spans 128 rows: the tail aliases later shared memory
                                           // (finite or not, those accumulator rows are never read)
constexpr uint32_t A_STAGE = 2 * A_PLANE;  // 6 KB, 1024-aligned

struct DtGeom {
  int I, E, C, H, W, HW;
  int h2, h4, h34, w2, w4, w34;
  int linear;
  int ntiles;
  const int* seg;                       // [I + 1]: expressions of image i are seg[i] .. seg[i + 1] - 1
};

struct DtMaps {
  CUtensorMap x, a_hi, a_lo;
};

// (hi, lo) bf16 pairs of two floats, packed for the K-major operand (element 0 in the low half)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hu << 16), rb = b - __uint_as_float(hu & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  *hi = hu;
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 (rows, cols) -> bf16 hi / lo planes, same layout ; block 0 also turns the non-decreasing expr2img into the
// per-image ranges seg[i] = first expression e with expr2img[e] >= i
__global__ void dt_split_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int n,
                                const int* __restrict__ e2i, int* __restrict__ seg, int E, int I) {
  if (blockIdx.x == 0)
    for (int e = threadIdx.x; e <= E; e += blockDim.x) {
      const int prev = e > 0 ? min(max(__ldg(e2i + e - 1), -1), I - 1) : -1;
      const int cur = e < E ? min(max(__ldg(e2i + e), -1), I - 1) : I;
      for (int i = prev + 1; i <= cur; ++i) seg[i] = e;
    }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = __ldg(src + i);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = __bfloat16_as_ushort(h);
    lo[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
  }
}

// loss[e] = sum_tiles partial[e][tile] / HW   (fixed order)
__global__ void dt_loss_kernel(const float* __restrict__ partial, float* __restrict__ loss, int E, int ntiles, float inv_hw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float s = 0.f;
  for (int t = 0; t < ntiles; ++t) s += partial[(size_t)e * ntiles + t];
  loss[e] = s * inv_hw;
}

template <int TPX>
struct DtPlan {
  static constexpr int DNST = TPX == 32 ? 6 : 4;               // ring stages
  static constexpr uint32_t B_PLANE = TPX * 128;               // TPX pixel rows x 128 B
  static constexpr uint32_t B_STAGE = 2 * B_PLANE;             // [X_hi rows ; X_lo rows]: one N = 2 TPX operand
  static constexpr int RAW_ROWS = 2 * ABOX;                    // drained accumulator rows: hi block, lo block
  static constexpr uint32_t RAW = RAW_ROWS * TPX * 4;
  static constexpr uint32_t GATE = DMAXE * TPX * 4;
  static size_t smem(int C) {
    return 1024 + DNST * (A_STAGE + B_STAGE) + (size_t)C * TPX * 4 + RAW + GATE + 256;
  }
};

template <int TPX>
__global__ void __launch_bounds__(DTHREADS, TPX == 32 ? 1 : 2)
dynfilter_tc_fwd_kernel(const __grid_constant__ DtMaps maps, const float* __restrict__ fuse, const int* __restrict__ e2i,
                        float* __restrict__ response, float* __restrict__ rk_saved, float* __restrict__ Y,
                        const float* __restrict__ target, float* __restrict__ loss_partial, DtGeom g) {
  using P = DtPlan<TPX>;
  constexpr int DNST = P::DNST;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  // layout: [A ring][B ring][X tile][drained accumulators][gates][barriers]
  uint8_t* a_ring = sm;                                         // DNST * A_STAGE
  uint8_t* b_ring = a_ring + DNST * A_STAGE;                    // DNST * B_STAGE
  float* xs = reinterpret_cast<float*>(b_ring + DNST * P::B_STAGE);   // [C][TPX]
  float* s_raw = xs + (size_t)g.C * TPX;                        // [2 ABOX][TPX], column rotated by the row
  float* s_gate = s_raw + P::RAW_ROWS * TPX;                    // [DMAXE][TPX]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_gate + DMAXE * TPX);
  uint64_t* xfull = bars;                // [4]
  uint64_t* full = bars + 4;             // [DNST]  A by TMA (transaction bytes) + one arrive per converter warp
  uint64_t* empty = full + DNST;         // [DNST]  tcgen05.commit
  uint64_t* tfull = empty + DNST;        // [1]
  static_assert((4 + 2 * DNST + 1) * 8 + 4 <= 256, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int img = blockIdx.y, p0 = blockIdx.x * TPX;
  // expressions of this image (expr2img is non-decreasing; ranges precomputed by dt_split_kernel)
  const int e0 = __ldg(g.seg + img), e1 = __ldg(g.seg + img + 1);
  if (e1 == e0) return;                  // block uniform
  const int nkb = g.C / DBK;
  const int nbox = (g.C + XBOX - 1) / XBOX;

  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&xfull[i], 1);
    for (int s = 0; s < DNST; ++s) {
      mbar_init(&full[s], 1 + 4);        // producer's arrive.expect_tx + one arrive per converter warp
      mbar_init(&empty[s], 1);           // tcgen05.commit
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * TPX);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {          // the X tile: once per CTA
    tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.a_hi); tc::prefetch_tmap(&maps.a_lo);
    for (int bx = 0; bx < nbox; ++bx) {
      const int rows = min(XBOX, g.C - bx * XBOX);
      mbar_arrive_expect_tx(&xfull[bx], (uint32_t)rows * TPX * 4u);
      tc::tma_load_2d(xs + (size_t)bx * XBOX * TPX, &maps.x, p0, img * g.C + bx * XBOX, &xfull[bx]);
    }
  }

  // expressions of the image in chunks of DMAXE (21 of the 24 box rows); ring positions run on across chunks
  int chunk = 0;
  for (int ec = e0; ec < e1; ec += DMAXE, ++chunk) {
    const int ne = min(DMAXE, e1 - ec);
    const int it0 = chunk * nkb;
    if (warp == 0) {
      // ================= TMA producer: filter k-blocks of this chunk through the ring =================
      if (lane == 0) {
        for (int kb = 0; kb < nkb; ++kb) {
          const int it = it0 + kb, s = it % DNST;
          if (it >= DNST) mbar_wait(&empty[s], ((it / DNST) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], A_STAGE);
          tc::tma_load_2d(a_ring + s * A_STAGE, &maps.a_hi, kb * DBK, ec * NF, &full[s]);
          tc::tma_load_2d(a_ring + s * A_STAGE + A_PLANE, &maps.a_lo, kb * DBK, ec * NF, &full[s]);
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = tc::make_idesc(2 * TPX, false, false);     // N = [X_hi ; X_lo] rows
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb, s = it % DNST;
        mbar_wait(&full[s], (it / DNST) & 1);
        tc::fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = base + s * A_STAGE, sb = base + DNST * A_STAGE + s * P::B_STAGE;
#pragma unroll
          for (int kk = 0; kk < DBK / 16; ++kk) {
            // rows 0..23 of A = f_hi, 24..47 = f_lo ; rows 0..TPX-1 of B = X_hi, TPX..2TPX-1 = X_lo
            const uint64_t a = tc::make_desc(sa + kk * 32, 0, 1024, tc::kLayoutSW128);
            const uint64_t b = tc::make_desc(sb + kk * 32, 0, 1024, tc::kLayoutSW128);
            tc::umma_f16(tmem_base, a, b, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty[s]);
          if (kb == nkb - 1) tc::umma_commit(tfull);
        }
        __syncwarp();
      }
    } else {
      // ================= converters: X tile fp32 -> (hi, lo) bf16 stages in the swizzled K-major layout =================
      // two groups of four warps take alternate stages; item = (pixel row of B, 16-byte chunk = 8 channels of the
      // 128-byte row): TPX * 8 items per stage over the 128 threads of a group
      const int grp = (warp - 2) >> 2;             // 0: warps 2-5, 1: warps 6-9
      const int ct = (t - 64) & 127;
      constexpr int ITEMS = TPX * 8 / 128;         // 2 (TPX = 32) or 1 (TPX = 16)
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb, s = it % DNST;
        if ((it & 1) != grp) continue;
        if (chunk == 0 && (kb % (XBOX / DBK)) < 2) mbar_wait(&xfull[(kb * DBK) / XBOX], 0);   // first touch of the box by this group
        if (it >= DNST) mbar_wait(&empty[s], ((it / DNST) - 1) & 1);
#pragma unroll
        for (int q = 0; q < ITEMS; ++q) {
          const int item = ct + q * 128;
          const int px = item % TPX, chk = item / TPX;
          const float* col = xs + (size_t)(kb * DBK + chk * 8) * TPX + px;
          uint32_t h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_pair(col[(2 * j) * TPX], col[(2 * j + 1) * TPX], &h[j], &l[j]);
          uint8_t* bs = b_ring + s * P::B_STAGE + px * 128 + ((chk ^ (px & 7)) << 4);
          *reinterpret_cast<uint4*>(bs) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(bs + P::B_PLANE) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
      // ================= drain: TMEM lane = stacked filter row; column blocks [X_hi | X_lo] are added up =================
      if (warp == 4 || warp == 5) {                // TMEM lane quadrants 0 (rows 0..31) and 1 (rows 32..63)
        const int q = warp & 3;
        mbar_wait(tfull, chunk & 1);
        tc::fence_after_sync();
        float v[TPX];
        if constexpr (TPX == 32) {
          float v2[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 32, v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        } else {
          float v2[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v2);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v2[j] + v2[16 + j];
        }
        const int row = q * 32 + lane;
        if (row < P::RAW_ROWS) {
#pragma unroll
          for (int j = 0; j < TPX; ++j) s_raw[row * TPX + ((j + row) & (TPX - 1))] = v[j];   // rotated: conflict free
        }
        tc::fence_before_sync();
      }
    }
    __syncthreads();
    // ---- thread = (expression, pixel): masks, r_k, r = sum_k w_k r_k, response, BCE partial, gate
    if (t < ((ne * TPX + 31) & ~31)) {             // whole warps: the loss reduction shuffles
      const bool act = t < ne * TPX;
      const int e = act ? t / TPX : 0, j = t % TPX;
      const int p = p0 + j;
      const int y = p / g.W, x = p - y * g.W;
      const bool m[NF] = {true, y < g.h2, y >= g.h2, x < g.w2, x >= g.w2, y >= g.h4 && y < g.h34, x >= g.w4 && x < g.w34};
      const bool in = act && p < g.HW;
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const int row = e * NF + k, row2 = row + ABOX;       // f_hi and f_lo blocks of the stacked operand
        const float rk = (in && m[k]) ? s_raw[row * TPX + ((j + row) & (TPX - 1))] + s_raw[row2 * TPX + ((j + row2) & (TPX - 1))]
                                      : 0.f;
        if (rk_saved && in) rk_saved[((size_t)(ec + e) * NF + k) * g.HW + p] = rk;
        r = fmaf(__ldg(fuse + (size_t)(ec + e) * NF + k), rk, r);
      }
      float l = 0.f;
      if (in) {
        response[(size_t)(ec + e) * g.HW + p] = r;
        if (target != nullptr) {
          const float tg = __ldg(target + (size_t)(ec + e) * g.HW + p);
          l = fmaxf(r, 0.f) - r * tg + log1pf(expf(-fabsf(r)));
        }
      }
      if (act) s_gate[e * TPX + j] = g.linear ? r : sigmoidf_acc(r);
      if (loss_partial != nullptr) {
#pragma unroll
        for (int o = TPX / 2; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o, TPX);   // within the TPX lanes of e
        if (act && j == 0) loss_partial[(size_t)(ec + e) * g.ntiles + blockIdx.x] = l;
      }
    }
    __syncthreads();
    // ---- gating: Y_e[c, tile] = X[c, tile] * gate_e ; thread = (pixel quad, channel slot)
    if (chunk == 0)
      for (int bx = 0; bx < nbox; ++bx) mbar_wait(&xfull[bx], 0);      // every reader of the tile observes the TMA completion
    {
      constexpr int QN = TPX / 4;                  // quads per tile row
      const int pq = t % QN, cs = t / QN;
      const int p = p0 + 4 * pq;
      if (p < g.HW) {                              // HW % 4 == 0: a quad is in or out as a whole
        float4 gq[DMAXE];
#pragma unroll
        for (int j = 0; j < DMAXE; ++j)
          gq[j] = (j < ne) ? reinterpret_cast<const float4*>(s_gate + j * TPX)[pq] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = cs; c < g.C; c += DTHREADS / QN) {
          const float4 x = reinterpret_cast<const float4*>(xs + (size_t)c * TPX)[pq];
#pragma unroll
          for (int j = 0; j < DMAXE; ++j)
            if (j < ne) {
              float4 y;
              y.x = x.x * gq[j].x; y.y = x.y * gq[j].y; y.z = x.z * gq[j].z; y.w = x.w * gq[j].w;
#if defined(L2S_DT_STORE) && L2S_DT_STORE == 1
              *reinterpret_cast<float4*>(Y + ((size_t)(ec + j) * g.C + c) * g.HW + p) = y;      // A/B: default write-back store
#elif defined(L2S_DT_STORE) && L2S_DT_STORE == 2
              __stwt(reinterpret_cast<float4*>(Y + ((size_t)(ec + j) * g.C + c) * g.HW + p), y);
#else
              __stcs(reinterpret_cast<float4*>(Y + ((size_t)(ec + j) * g.C + c) * g.HW + p), y);
#endif
            }
        }
      }
    }
    __syncthreads();                               // s_raw / s_gate are rewritten by the next chunk
  }
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 2 * TPX);
  }
}

// pixel tile: 32 (one CTA per SM, 128-byte rows) or 16 (two CTAs per SM); L2S_DYNFILTER_TPX overrides the default
int dt_tile() {
  static const int tpx = [] {
    const char* e = getenv("L2S_DYNFILTER_TPX");
    return (e && atoi(e) == 32) ? 32 : ((e && atoi(e) == 16) ? 16 : 16);
  }();
  return tpx;
}

int make_filter_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t C) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {DBK, ABOX}, estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  L2S_REQUIRE(r == CUDA_SUCCESS, L2S_ERR_CUDA, "dynfilter: cuTensorMapEncodeTiled(filters) failed with %d", (int)r);
  return L2S_OK;
}

template <int TPX>
int launch_tile(const DtMaps& maps, const float* fuse, const int* e2i, float* response, float* rk_saved, float* Y,
                const float* target, float* partial, const DtGeom& g, cudaStream_t st) {
  const size_t smem = DtPlan<TPX>::smem(g.C);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {      // once per device: the attribute is sticky
    L2S_CUDA_OK(cudaFuncSetAttribute(dynfilter_tc_fwd_kernel<TPX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)DtPlan<TPX>::smem(1024)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  dim3 grid(g.ntiles, g.I);
  dynfilter_tc_fwd_kernel<TPX><<<grid, DTHREADS, smem, st>>>(maps, fuse, e2i, response, rk_saved, Y, target, partial, g);
  L2S_LAUNCH_OK("dynfilter_tc_fwd_kernel");
  return L2S_OK;
}

}  // namespace

size_t dynfilter_tc_workspace_bytes(int E, int C, int HW) {
  const size_t plane = (((size_t)E * NF * C * 2) + 255) & ~(size_t)255;
  const size_t ntiles = (size_t)(HW + 15) / 16;                 // sized for the smaller tile
  return 2 * plane + (size_t)E * ntiles * sizeof(float) + 256 + 4096 * sizeof(int);   // + seg[I + 1], I <= 4095
}

// returns L2S_OK when the tensor-core kernel ran, 1 when the shape is outside its range (caller falls back), < 0 on error
int launch_dynfilter_tc_fwd(const float* X, const float* filt, const float* fuse, const int* e2i, float* response,
                            float* rk_saved, float* Y, const float* target, float* loss, int I, int E, int C, int H,
                            int W, int flags, void* workspace, size_t ws_bytes, cudaStream_t st) {
  const int HW = H * W;
  static const bool off = env_flag("L2S_DYNFILTER_FFMA");       // diagnostics / A-B: force the FFMA kernel
  const int tpx = dt_tile();
  const size_t smem = tpx == 32 ? DtPlan<32>::smem(C) : DtPlan<16>::smem(C);
  if (off || HW % 4 != 0 || C % DBK != 0 || C > 1024 || C < DBK || I > 4095 || (C > XBOX && C % XBOX != 0) || !workspace ||
      ws_bytes < dynfilter_tc_workspace_bytes(E, C, HW) || !aligned16(X) || !aligned16(Y) || !aligned16(workspace) ||
      (rk_saved && !aligned16(rk_saved)) || smem > (size_t)max_smem_optin())
    return 1;
  tc::EncodeTiledFn fn = tc::encode_fn();
  L2S_REQUIRE(fn, L2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  DtGeom g;
  g.I = I; g.E = E; g.C = C; g.H = H; g.W = W; g.HW = HW;
  g.h2 = H / 2; g.h4 = H / 4; g.h34 = (H * 3) / 4;
  g.w2 = W / 2; g.w4 = W / 4; g.w34 = (W * 3) / 4;
  g.linear = (flags & L2S_GATE_LINEAR) ? 1 : 0;
  g.ntiles = (HW + tpx - 1) / tpx;
  const size_t plane = (((size_t)E * NF * C * 2) + 255) & ~(size_t)255;
  uint16_t* a_hi = reinterpret_cast<uint16_t*>(workspace);
  uint16_t* a_lo = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(workspace) + plane);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 2 * plane);
  int* seg = reinterpret_cast<int*>(partial + (size_t)E * ((HW + 15) / 16) + 64);
  g.seg = seg;
  {
    const int n = E * NF * C;
    dt_split_kernel<<<min((n + 255) / 256, 4 * sm_count()), 256, 0, st>>>(filt, a_hi, a_lo, n, e2i, seg, E, I);
    L2S_LAUNCH_OK("dt_split_kernel");
  }
  DtMaps maps;
  {
    cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)I * C};
    cuuint64_t strides[1] = {(cuuint64_t)HW * 4};
    cuuint32_t box[2] = {(cuuint32_t)tpx, (cuuint32_t)(C < XBOX ? C : XBOX)}, estr[2] = {1, 1};
    CUresult r = fn(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    L2S_REQUIRE(r == CUDA_SUCCESS, L2S_ERR_CUDA, "dynfilter: cuTensorMapEncodeTiled(X) failed with %d", (int)r);
  }
  int rc;
  if ((rc = make_filter_map(&maps.a_hi, a_hi, (int64_t)E * NF, C))) return rc;
  if ((rc = make_filter_map(&maps.a_lo, a_lo, (int64_t)E * NF, C))) return rc;
  float* part = (loss && target) ? partial : nullptr;
  rc = tpx == 32 ? launch_tile<32>(maps, fuse, e2i, response, rk_saved, Y, target, part, g, st)
                 : launch_tile<16>(maps, fuse, e2i, response, rk_saved, Y, target, part, g, st);
  if (rc) return rc;
  if (loss && target) {
    dt_loss_kernel<<<(E + 127) / 128, 128, 0, st>>>(partial, loss, E, g.ntiles, 1.f / (float)HW);
    L2S_LAUNCH_OK("dt_loss_kernel");
    count_launch();
  } else if (loss) {
    L2S_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float) * E, st));
  }
  count_launch(2);
  return L2S_OK;
}

}  // namespace l2s

// Spatial dynamic-filter response layer, forward, on the Blackwell tensor pipe (north_star kernel (1)).
//
// Semantics: network_cycle_response.py:534-570 + response loss :415-422, identical to
// dynfilter.cu:  r_k[p] = M_k[p] sum_c f_k[c] X[c,p] ;  r = sum_k w_k r_k ;  Y = X * sigmoid(r).
//
// CTA = (image, tile of 32 pixels), 8 warps:
//   * the [C x 32 px] fp32 tile of X is brought in ONCE by TMA (2-D tensor map over (HW, I*C), four boxes of 256
//     channels, 128-byte rows = whole DRAM lines) and stays in shared memory: the contraction AND the gating read it
//     from there, so X costs its algorithmic bytes and nothing else;
//   * the contraction  D[(e,k), p] = sum_c f_{e,k}[c] X[c,p]  for ALL expressions of the image at once is a
//     [7 n_e (<= 21) -> 128] x [32 px] x [C] GEMM on tcgen05.mma (kind::f16, fp32 accumulators in TMEM) with the
//     bf16x3 split (hi*hi + hi*lo + lo*hi: ~2^-16 relative, inside the 1e-4 fp32 contract that rules out one-pass
//     TF32/bf16).  The stacked filter block arrives as bf16 (hi, lo) planes by TMA (64B-swizzled K-major boxes); the X
//     operand is produced from the resident fp32 tile by four converter warps, which write the (hi, lo) planes of a
//     32-channel k-block straight into the swizzled UMMA layout (one 16-byte chunk per thread and plane) and hand it to
//     the MMA warp through an mbarrier ring;
//   * epilogue out of TMEM: lane = filter row, 32 pixel columns per lane: partition mask, r_k (kept for the backward),
//     fuse weights -> r, sigmoid, BCE-with-logits partial per (expression, tile) (summed in fixed order by a second
//     tiny kernel: deterministic, no float atomics) -> gate;
//   * gating: Y_e = X * gate_e streamed out with float4 st.cs for every expression of the image.
// The filter block of a chunk is a 24-row TMA box (3 expressions x 7 filters; 128 rows would not fit next to the
// resident 128 KB tile), so an image with more than 3 expressions is processed in chunks of 3 against the same
// resident tile (the reference's batches carry 1-3 expressions per image).  Requirements (otherwise dynfilter.cu's FFMA kernel runs): H*W % 4 == 0 (TMA row stride),
// C % 32 == 0, C <= 1024 (tile residency), C <= 256 or C % 256 == 0 (whole TMA boxes of X).
#include <cuda.h>

#include "gemm_tc.cuh"

namespace l2s {
namespace {

constexpr int NF = L2S_NUM_FILTERS;
constexpr int TPX = 32;                 // pixels per CTA
constexpr int DBK = 32;                 // channels per k-block (64-byte bf16 rows)
constexpr int DNST = 4;                 // ring stages
constexpr int DTHREADS = 256;
constexpr int ABOX = 24;                // filter rows per TMA box
constexpr int DMAXE = ABOX / NF;        // expressions per chunk: 7 * 3 = 21 <= 24 box rows
constexpr int XBOX = 256;               // channels per TMA box of X
constexpr uint32_t A_PLANE = ABOX * 64;   // bytes per A plane of a stage actually filled by TMA (24 rows x 64 B); the
                                        // descriptor spans 128 rows: the tail aliases later shared memory (finite or
                                        // not, those accumulator rows are never read)
constexpr uint32_t A_STAGE = 2 * A_PLANE;
constexpr uint32_t B_PLANE = TPX * 64;  // 32 pixel rows x 64 B
constexpr uint32_t B_STAGE = 2 * B_PLANE;

struct DtGeom {
  int I, E, C, H, W, HW;
  int h2, h4, h34, w2, w4, w34;
  int linear;
  int ntiles;
};

struct DtMaps {
  CUtensorMap x, a_hi, a_lo;
};

__device__ __forceinline__ float mask_of(const DtGeom& g, int k, int p) {
  const int y = p / g.W, x = p - y * g.W;
  bool m;
  switch (k) {
    case 0: m = true; break;
    case 1: m = y < g.h2; break;
    case 2: m = y >= g.h2; break;
    case 3: m = x < g.w2; break;
    case 4: m = x >= g.w2; break;
    case 5: m = y >= g.h4 && y < g.h34; break;
    default: m = x >= g.w4 && x < g.w34; break;
  }
  return m ? 1.f : 0.f;
}

__device__ __forceinline__ void split_pair(float a, float b, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
  *hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  *lo = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
}

// fp32 (rows, cols) -> bf16 hi / lo planes, same layout
__global__ void dt_split_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int n) {
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

__global__ void __launch_bounds__(DTHREADS, 1)
dynfilter_tc_fwd_kernel(const __grid_constant__ DtMaps maps, const float* __restrict__ fuse, const int* __restrict__ e2i,
                        float* __restrict__ response, float* __restrict__ rk_saved, float* __restrict__ Y,
                        const float* __restrict__ target, float* __restrict__ loss_partial, DtGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  // layout: [A ring][B ring][X tile][epilogue scratch][barriers]
  uint8_t* a_ring = sm;                                         // DNST * A_STAGE (3 KB each; 1024-aligned stages)
  constexpr uint32_t A_STRIDE = 4096;                           // stage stride keeps every plane 512-byte aligned
  uint8_t* b_ring = a_ring + DNST * A_STRIDE;                   // DNST * B_STAGE
  float* xs = reinterpret_cast<float*>(b_ring + DNST * B_STAGE);   // [C][32]
  float* s_rk = xs + (size_t)g.C * TPX;                         // [32][32] masked, fuse-weighted r_k (rows < 21 used)
  float* s_gate = s_rk + 32 * TPX;                             // [DMAXE][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_gate + DMAXE * TPX);
  uint64_t* xfull = bars;                // [4]
  uint64_t* afull = bars + 4;            // [DNST]
  uint64_t* bfull = afull + DNST;        // [DNST]
  uint64_t* empty = bfull + DNST;        // [DNST]
  uint64_t* tfull = empty + DNST;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int img = blockIdx.y, p0 = blockIdx.x * TPX;
  // expressions of this image (expr2img is non-decreasing)
  int e0, e1;
  {
    int lo = 0, hi = g.E;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (__ldg(e2i + m) < img) lo = m + 1; else hi = m; }
    e0 = lo;
    hi = g.E;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (__ldg(e2i + m) <= img) lo = m + 1; else hi = m; }
    e1 = lo;
  }
  if (e1 == e0) return;                  // block uniform
  const int nkb = g.C / DBK;
  const int nbox = (g.C + XBOX - 1) / XBOX;

  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&xfull[i], 1);
    for (int s = 0; s < DNST; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&bfull[s], 4);           // one arrive per converter warp
      mbar_init(&empty[s], 1);           // tcgen05.commit
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 32);
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
          mbar_arrive_expect_tx(&afull[s], A_STAGE);
          tc::tma_load_2d(a_ring + s * A_STRIDE, &maps.a_hi, kb * DBK, ec * NF, &afull[s]);
          tc::tma_load_2d(a_ring + s * A_STRIDE + A_PLANE, &maps.a_lo, kb * DBK, ec * NF, &afull[s]);
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = tc::make_idesc(TPX, false, false);
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb, s = it % DNST;
        mbar_wait(&afull[s], (it / DNST) & 1);
        mbar_wait(&bfull[s], (it / DNST) & 1);
        tc::fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = base + s * A_STRIDE, sb = base + DNST * A_STRIDE + s * B_STAGE;
#pragma unroll
          for (int kk = 0; kk < DBK / 16; ++kk) {
            const uint64_t ahi = tc::make_desc(sa + kk * 32, 0, 512, tc::kLayoutSW64);
            const uint64_t alo = tc::make_desc(sa + A_PLANE + kk * 32, 0, 512, tc::kLayoutSW64);
            const uint64_t bhi = tc::make_desc(sb + kk * 32, 0, 512, tc::kLayoutSW64);
            const uint64_t blo = tc::make_desc(sb + B_PLANE + kk * 32, 0, 512, tc::kLayoutSW64);
            tc::umma_f16(tmem_base, alo, bhi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            tc::umma_f16(tmem_base, ahi, blo, idesc, 1u);
            tc::umma_f16(tmem_base, ahi, bhi, idesc, 1u);
          }
          tc::umma_commit(&empty[s]);
          if (kb == nkb - 1) tc::umma_commit(tfull);
        }
        __syncwarp();
      }
    } else if (warp < 6) {
      // ================= converters: X tile fp32 -> (hi, lo) bf16 k-blocks in the swizzled K-major layout =================
      const int ct = t - 64;                       // 0..127
      const int px = ct & 31, chk = ct >> 5;       // pixel row of B, 16-byte chunk (8 channels) of the 64-byte row
      const uint32_t phys = (uint32_t)(chk ^ ((px >> 1) & 3));
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb, s = it % DNST;
        if (chunk == 0 && (kb * DBK) % XBOX == 0) mbar_wait(&xfull[(kb * DBK) / XBOX], 0);
        if (it >= DNST) mbar_wait(&empty[s], ((it / DNST) - 1) & 1);
        const float* col = xs + (size_t)(kb * DBK + chk * 8) * TPX + px;
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_pair(col[(2 * j) * TPX], col[(2 * j + 1) * TPX], &h[j], &l[j]);
        uint8_t* bs = b_ring + s * B_STAGE + px * 64 + phys * 16;
        *reinterpret_cast<uint4*>(bs) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(bs + B_PLANE) = make_uint4(l[0], l[1], l[2], l[3]);
        fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&bfull[s]);
      }
      // ================= epilogue: TMEM lane = filter row (e, k), 32 pixel columns =================
      mbar_wait(tfull, chunk & 1);
      tc::fence_after_sync();
      const int q = warp & 3;                      // TMEM lane quadrant of this warp
      const int row = q * 32 + lane;
      if (q * 32 < ne * NF) {                      // warp uniform: quadrants without real rows skip the load
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
        if (row < ne * NF) {
          const int e = row / NF, k = row - e * NF;
          const float w = __ldg(fuse + (size_t)(ec + e) * NF + k);
          float* rk_out = rk_saved ? rk_saved + ((size_t)(ec + e) * NF + k) * g.HW + p0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int p = p0 + j;
            const float rk = (p < g.HW) ? mask_of(g, k, p) * v[j] : 0.f;
            v[j] = rk;
            s_rk[row * TPX + ((j + row) & 31)] = w * rk;        // rotated: conflict-free for the row-per-lane writes
          }
          if (rk_out) {
            if (p0 + TPX <= g.HW) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(rk_out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (p0 + j < g.HW) rk_out[j] = v[j];
            }
          }
        }
      }
      tc::fence_before_sync();
    } else if (chunk == 0) {
      // warps 6, 7 only gate: they need the whole X tile
      for (int bx = 0; bx < nbox; ++bx) mbar_wait(&xfull[bx], 0);
    }
    __syncthreads();
    // ---- r = sum_k w_k r_k, gate, response, BCE partial: warp w takes expressions w, w + 8, ...; lane = pixel
    for (int e = warp; e < ne; e += DTHREADS / 32) {
      const int p = p0 + lane;
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const int row = e * NF + k;
        r += s_rk[row * TPX + ((lane + row) & 31)];
      }
      float l = 0.f;
      if (p < g.HW) {
        response[(size_t)(ec + e) * g.HW + p] = r;
        if (target != nullptr) {
          const float tg = __ldg(target + (size_t)(ec + e) * g.HW + p);
          l = fmaxf(r, 0.f) - r * tg + log1pf(expf(-fabsf(r)));
        }
      }
      s_gate[e * TPX + lane] = g.linear ? r : sigmoidf_acc(r);
      if (loss_partial != nullptr) {
        l = warp_sum(l);
        if (lane == 0) loss_partial[(size_t)(ec + e) * g.ntiles + blockIdx.x] = l;
      }
    }
    __syncthreads();
    // ---- gating: Y_e[c, tile] = X[c, tile] * gate_e ; thread = (pixel quad, channel slot), whole 128-byte rows per warp
    if (chunk == 0)
      for (int bx = 0; bx < nbox; ++bx) mbar_wait(&xfull[bx], 0);      // every reader of the tile observes the TMA completion
    {
      const int pq = t & 7, cs = t >> 3;           // 8 quads x 32 channel slots
      const int p = p0 + 4 * pq;
      if (p < g.HW) {                              // HW % 4 == 0: a quad is in or out as a whole
        for (int eb = 0; eb < ne; eb += 3) {
          const int nb = min(3, ne - eb);
          float4 gq[3];
#pragma unroll
          for (int j = 0; j < 3; ++j)
            gq[j] = (j < nb) ? reinterpret_cast<const float4*>(s_gate + (eb + j) * TPX)[pq] : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int c = cs; c < g.C; c += DTHREADS / 8) {
            const float4 x = reinterpret_cast<const float4*>(xs + (size_t)c * TPX)[pq];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (j < nb) {
                float4 y;
                y.x = x.x * gq[j].x; y.y = x.y * gq[j].y; y.z = x.z * gq[j].z; y.w = x.w * gq[j].w;
                __stcs(reinterpret_cast<float4*>(Y + ((size_t)(ec + eb + j) * g.C + c) * g.HW + p), y);
              }
          }
        }
      }
    }
    __syncthreads();                               // s_rk / s_gate are rewritten by the next chunk
  }
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 32);
  }
}

size_t dt_smem(int C) {
  return 1024 + DNST * 4096 + DNST * B_STAGE + (size_t)C * TPX * 4 + 32 * TPX * 4 + DMAXE * TPX * 4 + 256;
}

}  // namespace

size_t dynfilter_tc_workspace_bytes(int E, int C, int HW) {
  const size_t plane = (((size_t)E * NF * C * 2) + 255) & ~(size_t)255;
  const size_t ntiles = (size_t)(HW + TPX - 1) / TPX;
  return 2 * plane + (size_t)E * ntiles * sizeof(float) + 256;
}

// returns L2S_OK when the tensor-core kernel ran, 1 when the shape is outside its range (caller falls back), < 0 on error
int launch_dynfilter_tc_fwd(const float* X, const float* filt, const float* fuse, const int* e2i, float* response,
                            float* rk_saved, float* Y, const float* target, float* loss, int I, int E, int C, int H,
                            int W, int flags, void* workspace, size_t ws_bytes, cudaStream_t st) {
  const int HW = H * W;
  static const bool off = env_flag("L2S_DYNFILTER_FFMA");       // diagnostics / A-B: force the FFMA kernel
  if (off || HW % 4 != 0 || C % DBK != 0 || C > 1024 || C < DBK || (C > XBOX && C % XBOX != 0) || !workspace ||
      ws_bytes < dynfilter_tc_workspace_bytes(E, C, HW) || !aligned16(X) || !aligned16(Y) || !aligned16(workspace) ||
      (rk_saved && !aligned16(rk_saved)) || dt_smem(C) > (size_t)max_smem_optin())
    return 1;
  tc::EncodeTiledFn fn = tc::encode_fn();
  L2S_REQUIRE(fn, L2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  DtGeom g;
  g.I = I; g.E = E; g.C = C; g.H = H; g.W = W; g.HW = HW;
  g.h2 = H / 2; g.h4 = H / 4; g.h34 = (H * 3) / 4;
  g.w2 = W / 2; g.w4 = W / 4; g.w34 = (W * 3) / 4;
  g.linear = (flags & L2S_GATE_LINEAR) ? 1 : 0;
  g.ntiles = (HW + TPX - 1) / TPX;
  const size_t plane = (((size_t)E * NF * C * 2) + 255) & ~(size_t)255;
  uint16_t* a_hi = reinterpret_cast<uint16_t*>(workspace);
  uint16_t* a_lo = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(workspace) + plane);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 2 * plane);
  {
    const int n = E * NF * C;
    dt_split_kernel<<<min((n + 255) / 256, 4 * sm_count()), 256, 0, st>>>(filt, a_hi, a_lo, n);
    L2S_LAUNCH_OK("dt_split_kernel");
  }
  DtMaps maps;
  {
    cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)I * C};
    cuuint64_t strides[1] = {(cuuint64_t)HW * 4};
    cuuint32_t box[2] = {TPX, (cuuint32_t)(C < XBOX ? C : XBOX)}, estr[2] = {1, 1};
    CUresult r = fn(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    L2S_REQUIRE(r == CUDA_SUCCESS, L2S_ERR_CUDA, "dynfilter: cuTensorMapEncodeTiled(X) failed with %d", (int)r);
  }
  int rc;
  if ((rc = tc::make_operand_map(&maps.a_hi, a_hi, (int64_t)E * NF, C, C, false, ABOX))) return rc;
  if ((rc = tc::make_operand_map(&maps.a_lo, a_lo, (int64_t)E * NF, C, C, false, ABOX))) return rc;
  const size_t smem = dt_smem(C);
  L2S_CUDA_OK(cudaFuncSetAttribute(dynfilter_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(g.ntiles, I);
  dynfilter_tc_fwd_kernel<<<grid, DTHREADS, smem, st>>>(maps, fuse, e2i, response, rk_saved, Y, target,
                                                       (loss && target) ? partial : nullptr, g);
  L2S_LAUNCH_OK("dynfilter_tc_fwd_kernel");
  if (loss && target) {
    dt_loss_kernel<<<(E + 127) / 128, 128, 0, st>>>(partial, loss, E, g.ntiles, 1.f / (float)HW);
    L2S_LAUNCH_OK("dt_loss_kernel");
    count_launch();
  }
  count_launch(2);
  return L2S_OK;
}

}  // namespace l2s

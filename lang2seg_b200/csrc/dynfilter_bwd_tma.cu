// Spatial dynamic-filter response layer, backward, TMA-streamed (north_star kernel (1), backward half).
//
// Semantics: the backward of network_cycle_response.py:534-570 (+ response loss :415-422), identical to
// dynfilter_bwd_kernel in dynfilter.cu:
//   ds[p]   = sum_c dY_e[c,p] X[c,p]                     dr_e[p] = ds * g'(r) + dresponse + resp_gscale * (sigmoid(r) - t) / HW
//   dX[c,p] = sum_e ( dY_e[c,p] g_e[p] + sum_k f_{e,k}[c] w_{e,k} M_k[p] dr_e[p] )
// (dfilt / dfuse need dr over ALL pixels of the image and are produced from `drbuf` by the two small kernels that
// follow, dynfilter.cu.)
//
// CTA = (image, tile of 32 pixels), 16 compute warps + 2 producer warps:
//   * the [C x 32 px] tile of X arrives ONCE by TMA (the forward's tensor map: 128-byte rows) and stays in shared
//     memory; the gradient tiles of the image's expressions are STREAMED through a ring of [64 ch x 32 px] TMA boxes
//     (8 KB each, one mbarrier pair per box) by a producer warp that runs ahead across expression boundaries, so the
//     loads of expression e+1 are in flight while expression e is being reduced -- the FFMA tile kernel loaded with
//     8 x 16 bytes in flight per thread between block-wide barriers and reached 0.18 of the HBM roofline;
//   * thread = (pixel quad, channel slot): the dX tile lives in registers (16 x float4 per thread for C = 1024) for
//     the whole CTA -- no shared-memory accumulator, no read-modify-write traffic;
//   * per expression: one sweep over the ring (ds partials + the gate term), a fixed-order reduction of ds over the
//     channel slots (shuffles, then 16 per-warp partials in shared memory), dr / the 7 masked weights per pixel, and the
//     filter term as 7 FMAs per element against the expression's filters (28 KB, bulk-copied by the second producer
//     warp while the sweep runs).
// Deterministic (fixed summation order, no atomics).  Requirements (otherwise the kernels of dynfilter.cu run):
// H*W % 4 == 0, C % 64 == 0, C <= 1024, C <= 256 or C % 256 == 0.
#include <cuda.h>

#include "gemm_tc.cuh"

namespace l2s {
namespace {

constexpr int NF = L2S_NUM_FILTERS;
constexpr int TPX = 32;
constexpr int BCH = 64;                  // channels per ring box
constexpr int NST = 7;                   // ring stages (8 KB each)
constexpr int NCW = 16;                  // compute warps
constexpr int NTHREADS = (NCW + 2) * 32;
constexpr int MAXB = 1024 / BCH;         // boxes per expression at C = 1024
constexpr int XBOX = 256;
constexpr uint32_t BOX_BYTES = BCH * TPX * 4;

struct BtGeom {
  int I, E, C, H, W, HW;
  int h2, h4, h34, w2, w4, w34;
  int linear;
};

struct BtMaps {
  CUtensorMap x, dy;
};

// seg[i] = first expression e with expr2img[e] >= i  (expr2img non-decreasing)
__global__ void bt_seg_kernel(const int* __restrict__ e2i, int* __restrict__ seg, int E, int I) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e <= E; e += gridDim.x * blockDim.x) {
    const int prev = e > 0 ? min(max(__ldg(e2i + e - 1), -1), I - 1) : -1;
    const int cur = e < E ? min(max(__ldg(e2i + e), -1), I - 1) : I;
    for (int i = prev + 1; i <= cur; ++i) seg[i] = e;
  }
}

size_t bt_smem(int C) {
  return 1024 + (size_t)C * TPX * 4 + NST * BOX_BYTES + (size_t)NF * C * 4 + NCW * TPX * 4 + (NF + 2) * TPX * 4 + 256;
}

__global__ void __launch_bounds__(NTHREADS, 1)
dynfilter_bwd_tma_kernel(const __grid_constant__ BtMaps maps, const float* __restrict__ filt, const float* __restrict__ fuse,
                         const int* __restrict__ seg, const float* __restrict__ response, const float* __restrict__ dresp,
                         const float* __restrict__ target, const float* __restrict__ gscale, float* __restrict__ dX,
                         float* __restrict__ drbuf, BtGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* xs = reinterpret_cast<float*>(sm);                         // [C][32]
  float* ring = xs + (size_t)g.C * TPX;                             // [NST][64][32]
  float* fs = ring + NST * BCH * TPX;                               // [7][C]
  float* red = fs + (size_t)NF * g.C;                               // [NCW][32] ds partials
  float* s_gate = red + NCW * TPX;                                  // [32]
  float* s_dsum = s_gate + TPX;                                     // [32] (unused slot kept for alignment)
  float* mwdr = s_dsum + TPX;                                       // [7][32]  w_k M_k[p] dr[p]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mwdr + NF * TPX);
  uint64_t* xfull = bars;                // [4]
  uint64_t* full = bars + 4;             // [NST]
  uint64_t* empty = full + NST;          // [NST]  one arrive per compute warp
  uint64_t* fsfull = empty + NST;        // [1]
  uint64_t* fsfree = fsfull + 1;         // [1]    one arrive per compute warp after the filter term

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int img = blockIdx.y, p0 = blockIdx.x * TPX;
  const int e0 = __ldg(seg + img), e1 = __ldg(seg + img + 1);
  const int nb = g.C / BCH;
  const int nbox = (g.C + XBOX - 1) / XBOX;

  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&xfull[i], 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    mbar_init(fsfull, 1);
    mbar_init(fsfree, NCW);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NCW) {
    // ================= producer 1: the X tile, then the gradient boxes of every expression through the ring =================
    if (lane == 0 && e1 > e0) {
      tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.dy);
      for (int bx = 0; bx < nbox; ++bx) {
        const int rows = min(XBOX, g.C - bx * XBOX);
        mbar_arrive_expect_tx(&xfull[bx], (uint32_t)rows * TPX * 4u);
        tc::tma_load_2d(xs + (size_t)bx * XBOX * TPX, &maps.x, p0, img * g.C + bx * XBOX, &xfull[bx]);
      }
      int it = 0;
      for (int e = e0; e < e1; ++e)
        for (int b = 0; b < nb; ++b, ++it) {
          const int s = it % NST;
          if (it >= NST) mbar_wait(&empty[s], ((it / NST) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], BOX_BYTES);
          tc::tma_load_2d(ring + (size_t)s * BCH * TPX, &maps.dy, p0, e * g.C + b * BCH, &full[s]);
        }
    }
    return;
  }
  if (warp == NCW + 1) {
    // ================= producer 2: the filters of each expression (28 KB), as soon as the previous ones are done with ===
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(NF * g.C * sizeof(float));
      for (int e = e0, j = 0; e < e1; ++e, ++j) {
        if (j > 0) mbar_wait(fsfree, (j - 1) & 1);
        mbar_arrive_expect_tx(fsfull, bytes);
        bulk_g2s(fs, filt + (size_t)e * NF * g.C, bytes, fsfull);
      }
    }
    return;
  }

  // ================= compute warps: thread = (pixel quad pq, channel slot cs); channel of box b = 64 b + cs =================
  const int pq = t & 7, cs = t >> 3;
  const int p = p0 + 4 * pq;
  const bool pin = p < g.HW;                        // HW % 4 == 0: a quad is in or out as a whole
  float4 acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);

  int it = 0;
  for (int e = e0, j = 0; e < e1; ++e, ++j) {
    if (t < TPX) {
      const int pp = p0 + t;
      const float r = (pp < g.HW) ? __ldg(response + (size_t)e * g.HW + pp) : 0.f;
      s_gate[t] = g.linear ? r : sigmoidf_acc(r);
    }
    // named barrier over the compute warps only (the producers have left): gate visible, previous mwdr / red consumed
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    const float4 gq = reinterpret_cast<const float4*>(s_gate)[pq];
    if (j == 0)
      for (int bx = 0; bx < nbox; ++bx) mbar_wait(&xfull[bx], 0);
    // ---- sweep: ds partials and the gate term
    float4 dsp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < nb) {
        const int s = it % NST;
        if (lane == 0) mbar_wait(&full[s], (it / NST) & 1);       // one poller per warp
        __syncwarp();
        const float4 v = reinterpret_cast<const float4*>(ring + (size_t)s * BCH * TPX)[cs * (TPX / 4) + pq];
        const float4 x = reinterpret_cast<const float4*>(xs)[(size_t)(b * BCH + cs) * (TPX / 4) + pq];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        dsp.x = fmaf(v.x, x.x, dsp.x); dsp.y = fmaf(v.y, x.y, dsp.y);
        dsp.z = fmaf(v.z, x.z, dsp.z); dsp.w = fmaf(v.w, x.w, dsp.w);
        acc[b].x = fmaf(v.x, gq.x, acc[b].x); acc[b].y = fmaf(v.y, gq.y, acc[b].y);
        acc[b].z = fmaf(v.z, gq.z, acc[b].z); acc[b].w = fmaf(v.w, gq.w, acc[b].w);
        ++it;
      }
    }
    // ---- ds over the 64 channel slots: 4 slots of the warp by shuffles (lane bits 3, 4), 16 warps through shared memory
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      dsp.x += __shfl_xor_sync(0xffffffffu, dsp.x, o);
      dsp.y += __shfl_xor_sync(0xffffffffu, dsp.y, o);
      dsp.z += __shfl_xor_sync(0xffffffffu, dsp.z, o);
      dsp.w += __shfl_xor_sync(0xffffffffu, dsp.w, o);
    }
    if (lane < 8) reinterpret_cast<float4*>(red + warp * TPX)[lane] = dsp;
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    if (t < TPX) {
      const int pp = p0 + t;
      float ds = 0.f;
#pragma unroll
      for (int w = 0; w < NCW; ++w) ds += red[w * TPX + t];
      float dr = 0.f;
      if (pp < g.HW) {
        const float r = __ldg(response + (size_t)e * g.HW + pp);
        const float sg = sigmoidf_acc(r);
        dr = g.linear ? ds : ds * sg * (1.f - sg);
        if (dresp) dr += __ldg(dresp + (size_t)e * g.HW + pp);
        if (target && gscale) dr += __ldg(gscale + e) * (sg - __ldg(target + (size_t)e * g.HW + pp)) / (float)g.HW;
        drbuf[(size_t)e * g.HW + pp] = dr;
      }
      const int y = pp / g.W, x = pp - y * g.W;
      const bool m[NF] = {true, y < g.h2, y >= g.h2, x < g.w2, x >= g.w2, y >= g.h4 && y < g.h34, x >= g.w4 && x < g.w34};
#pragma unroll
      for (int k = 0; k < NF; ++k)
        mwdr[k * TPX + t] = (pp < g.HW && m[k]) ? __ldg(fuse + (size_t)e * NF + k) * dr : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    // ---- filter term: dX[c,p] += sum_k f_k[c] * (w_k M_k dr)[p]
    mbar_wait(fsfull, j & 1);
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      const float4 mk = reinterpret_cast<const float4*>(mwdr + k * TPX)[pq];
      const float* fk = fs + (size_t)k * g.C + cs;
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < nb) {
          const float f = fk[b * BCH];
          acc[b].x = fmaf(f, mk.x, acc[b].x); acc[b].y = fmaf(f, mk.y, acc[b].y);
          acc[b].z = fmaf(f, mk.z, acc[b].z); acc[b].w = fmaf(f, mk.w, acc[b].w);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(fsfree);
  }
  // ---- the dX tile leaves the registers (also for an image without expressions: zeros)
  if (pin) {
    float* dXi = dX + (size_t)img * g.C * g.HW + p;
#pragma unroll
    for (int b = 0; b < MAXB; ++b)
      if (b < nb) __stcs(reinterpret_cast<float4*>(dXi + (size_t)(b * BCH + cs) * g.HW), acc[b]);
  }
}

int make_f32_map(CUtensorMap* out, const float* ptr, int64_t rows, int HW, int box_rows) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  L2S_REQUIRE(fn, L2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)HW * 4};
  cuuint32_t box[2] = {TPX, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  L2S_REQUIRE(r == CUDA_SUCCESS, L2S_ERR_CUDA, "dynfilter_bwd: cuTensorMapEncodeTiled failed with %d", (int)r);
  return L2S_OK;
}

}  // namespace

size_t dynfilter_bwd_tma_workspace_bytes(int I) { return ((size_t)I + 2) * sizeof(int) + 256; }

// returns L2S_OK when the TMA-streamed kernel ran, 1 when the shape is outside its range (caller falls back), < 0 on error
int launch_dynfilter_bwd_tma(const float* X, const float* filt, const float* fuse, const int* e2i, const float* response,
                             const float* dY, const float* dresp, const float* target, const float* gscale, float* dX,
                             float* drbuf, int I, int E, int C, int H, int W, int flags, void* seg_ws, size_t seg_bytes,
                             cudaStream_t st) {
  const int HW = H * W;
  static const bool off = env_flag("L2S_DYNFILTER_BWD_FFMA");   // diagnostics / A-B: force the FFMA tile kernels
  if (off || HW % 4 != 0 || C % BCH != 0 || C > 1024 || (C > XBOX && C % XBOX != 0) || !seg_ws ||
      seg_bytes < dynfilter_bwd_tma_workspace_bytes(I) || !aligned16(X) || !aligned16(dY) || !aligned16(dX) ||
      !aligned16(filt) || (NF * C * 4) % 16 != 0 || bt_smem(C) > (size_t)max_smem_optin())
    return 1;
  BtGeom g;
  g.I = I; g.E = E; g.C = C; g.H = H; g.W = W; g.HW = HW;
  g.h2 = H / 2; g.h4 = H / 4; g.h34 = (H * 3) / 4;
  g.w2 = W / 2; g.w4 = W / 4; g.w34 = (W * 3) / 4;
  g.linear = (flags & L2S_GATE_LINEAR) ? 1 : 0;
  int* seg = reinterpret_cast<int*>(seg_ws);
  bt_seg_kernel<<<(E + 256) / 256, 256, 0, st>>>(e2i, seg, E, I);
  L2S_LAUNCH_OK("bt_seg_kernel");
  BtMaps maps;
  int rc;
  if ((rc = make_f32_map(&maps.x, X, (int64_t)I * C, HW, C < XBOX ? C : XBOX))) return rc;
  if ((rc = make_f32_map(&maps.dy, dY, (int64_t)E * C, HW, BCH))) return rc;
  const size_t smem = bt_smem(C);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    L2S_CUDA_OK(cudaFuncSetAttribute(dynfilter_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem(1024)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  dim3 grid((HW + TPX - 1) / TPX, I);
  dynfilter_bwd_tma_kernel<<<grid, NTHREADS, smem, st>>>(maps, filt, fuse, seg, response, dresp, target, gscale, dX, drbuf, g);
  L2S_LAUNCH_OK("dynfilter_bwd_tma_kernel");
  count_launch(2);
  return L2S_OK;
}

}  // namespace l2s

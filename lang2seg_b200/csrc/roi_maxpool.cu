// Caffe-style RoI max-pool forward / backward for sm_100a (the reference's POOLING_MODE != 'crop').
//
// Semantics: pyutils/mask-faster-rcnn/lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.cu
//   forward :15-75  -- round(roi*scale), float bin sizes, [floor, ceil) bins clipped to the map,
//                      first strict maximum in row-major order, argmax = flat (c*H+h)*W+w, empty
//                      bin -> 0 / -1.  Integer contract: bins and argmax are bit-exact.
//   backward:104-179 -- the reference gathers over all ROIs per input element (O(R) each); the
//                      same sums are produced here by an owner-computes scatter in shared memory
//                      (deterministic; roi_maxpool_bwd_owner_kernel), atomics only for huge maps.
// Unlike the reference (batch must be 1, roi_pooling_cuda.c:27-30) any batch size is accepted.
#include "common.cuh"

namespace l2s {
namespace {

__global__ void roi_maxpool_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                       float* __restrict__ out, int* __restrict__ argmax, int B, int C, int H,
                                       int W, int N, int PH, int PW, float scale) {
  const size_t total = (size_t)N * C * PH * PW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int pw = (int)(r % PW); r /= PW;
    const int ph = (int)(r % PH); r /= PH;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    const float* roi = rois + 5 * (size_t)n;
    const int b = (int)__ldg(roi);
    const int rsw = (int)roundf(__ldg(roi + 1) * scale);
    const int rsh = (int)roundf(__ldg(roi + 2) * scale);
    const int rew = (int)roundf(__ldg(roi + 3) * scale);
    const int reh = (int)roundf(__ldg(roi + 4) * scale);
    const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);   // malformed ROIs become 1x1
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
    int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + rsh, 0), H);
    he = min(max(he + rsh, 0), H);
    ws = min(max(ws + rsw, 0), W);
    we = min(max(we + rsw, 0), W);
    const bool empty = (he <= hs) || (we <= ws) || (unsigned)b >= (unsigned)B;
    float best = empty ? 0.f : -3.402823466e+38F;
    int bi = -1;
    if (!empty) {
      const float* f = feat + (size_t)b * C * H * W;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          const int fi = (c * H + h) * W + w;
          const float v = __ldg(f + fi);
          if (v > best) { best = v; bi = fi; }
        }
    }
    out[idx] = best;
    if (argmax) argmax[idx] = bi;
  }
}

__global__ void roi_maxpool_bwd_kernel(const float* __restrict__ top, const float* __restrict__ rois,
                                       const int* __restrict__ argmax, float* __restrict__ bottom, int B, int C,
                                       int H, int W, int N, int PH, int PW) {
  const size_t per = (size_t)C * PH * PW, total = (size_t)N * per;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / per);
    const int a = argmax[idx];
    const int b = (int)__ldg(rois + 5 * (size_t)n);
    if (a >= 0 && (unsigned)b < (unsigned)B) atomicAdd(bottom + (size_t)b * C * H * W + a, top[idx]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Deterministic backward: CTA = (image, chunk of CC channels) with the gradient of the map slice resident in shared
// memory as [pixel][CC+1] (lane = channel, odd stride).  The ROIs of the image are found by an order-preserving
// compaction of the batch-index column, 256 at a time.
// (Forward: two shared-memory-resident variants of this layout -- warp = bin with block barriers per ROI, 12.5 ms, and
// warp = ROI with bulk stores, 14.3 ms at cfg-2 sizes -- lost against the thread-per-output kernel above, 8.2 ms: with
// the 135 KB slice resident only 6-8 warps fit per SM and the bin scans are latency bound; the gather kernel hides
// the same latency behind 64 warps per SM out of L2.  Removed; numbers in profiles/r02_ab.md.)
// ---------------------------------------------------------------------------------------------------------------------
// ROIs of image b among base .. base + 255 -> s_list (ascending), returns their number.  All 256 threads call it.
__device__ __forceinline__ int compact_rois(const float* __restrict__ rois, int N, int base, int b, int* s_list, int* s_wcnt) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int n = base + t;
  const bool mine = n < N && (int)__ldg(rois + 5 * (size_t)n) == b;
  const unsigned bal = __ballot_sync(0xffffffffu, mine);
  __syncthreads();                       // the previous block's list and counts are no longer in use
  if (lane == 0) s_wcnt[wid] = __popc(bal);
  __syncthreads();
  int pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < wid) pre += s_wcnt[w];
    tot += s_wcnt[w];
  }
  if (mine) s_list[pre + __popc(bal & ((1u << lane) - 1u))] = n;
  __syncthreads();
  return tot;
}

// Deterministic owner-computes backward.  The gradient of the map slice accumulates in shared memory; the CTA walks the
// ROIs of its image in ascending order, the (top_grad, argmax) tiles of the next ROIs arrive by 1-D bulk copies (TMA engine,
// three in flight) while the current one is applied, and warp w (lane group) applies exactly the bins whose arg-max pixel
// lies in ITS contiguous pixel band -- every accumulator has one owner and is updated in a fixed (ROI, bin) order:
// bit-reproducible, no atomics (the scatter kernel above sums in arrival order).  The slice is written once at the end;
// no memset of the output is needed.
constexpr int kOwnStages = 3;
template <int CC>
__global__ void __launch_bounds__(256)
roi_maxpool_bwd_owner_kernel(const float* __restrict__ top, const float* __restrict__ rois, const int* __restrict__ argmax,
                             float* __restrict__ bottom, int B, int C, int H, int W, int N, int PP, int bulk) {
  constexpr int LD = CC + 1;
  extern __shared__ __align__(16) float smem[];
  const int HW = H * W;
  const int TILE = (CC * PP + 3) & ~3;                     // floats per staged tile (16-byte multiple)
  float* acc = smem;                                       // [HW][LD]
  float* s_top = acc + (((size_t)HW * LD + 3) & ~(size_t)3);   // [kOwnStages][TILE]
  int* s_arg = reinterpret_cast<int*>(s_top + kOwnStages * TILE);   // [kOwnStages][TILE]
  int* s_list = s_arg + kOwnStages * TILE;                 // [256] ROI ids of this image in the current block of 256
  __shared__ int s_wcnt[8];
  __shared__ uint64_t full[kOwnStages];
  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, C - c0);
  for (int i = t; i < HW * LD; i += 256) acc[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < kOwnStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  // owner = (warp, lane group): lane = channel, and with CC < 32 the 32 / CC lane groups of a warp own separate pixel
  // bands (never the same accumulator from two lanes of one instruction)
  constexpr int SUB = 32 / CC;
  const int ch = lane % CC, band = wid * SUB + lane / CC;
  const int lo = (int)(((long long)HW * band) / (8 * SUB)), hi = (int)(((long long)HW * (band + 1)) / (8 * SUB));
  const int coff = (c0 + ch) * HW;
  const uint32_t tile_bytes = (uint32_t)(cvalid * PP * 4);
  int it = 0;                                              // tiles consumed so far (ring position / parity)
  for (int base = 0; base < N; base += 256) {
    const int tot = compact_rois(rois, N, base, b, s_list, s_wcnt);
    auto issue = [&](int q, int slot) {                    // thread 0: both tiles of ROI s_list[q] into `slot`
      const size_t src = ((size_t)s_list[q] * C + c0) * PP;
      mbar_arrive_expect_tx(&full[slot], 2 * tile_bytes);
      bulk_g2s(s_top + slot * TILE, top + src, tile_bytes, &full[slot]);
      bulk_g2s(s_arg + slot * TILE, argmax + src, tile_bytes, &full[slot]);
    };
    if (bulk && t == 0)
      for (int q = 0; q < min(tot, kOwnStages); ++q) issue(q, (it + q) % kOwnStages);
    for (int q = 0; q < tot; ++q, ++it) {
      const int slot = it % kOwnStages;
      if (bulk) {
        mbar_wait(&full[slot], (it / kOwnStages) & 1);
      } else {
        const size_t src = ((size_t)s_list[q] * C + c0) * PP;
        for (int i = t; i < cvalid * PP; i += 256) {
          s_top[slot * TILE + i] = __ldg(top + src + i);
          s_arg[slot * TILE + i] = __ldg(argmax + src + i);
        }
        __syncthreads();
      }
      if (ch < cvalid) {
        const float* tt = s_top + slot * TILE + ch * PP;
        const int* aa = s_arg + slot * TILE + ch * PP;
        for (int bin = 0; bin < PP; ++bin) {
          const int px = aa[bin] - coff;                   // empty bins (-1) fall below every band
          if (px >= lo && px < hi) acc[(size_t)px * LD + ch] += tt[bin];
        }
      }
      __syncthreads();                                     // the slot has been read by everyone
      if (bulk && t == 0 && q + kOwnStages < tot) issue(q + kOwnStages, slot);
    }
  }
  __syncthreads();
  float* dst = bottom + ((size_t)b * C + c0) * HW;
  for (int i = t; i < cvalid * HW; i += 256) {
    const int c = i / HW, px = i - c * HW;
    dst[i] = acc[(size_t)px * LD + c];
  }
}

template <int CC>
size_t owner_smem(int HW, int PP) {
  return (((size_t)HW * (CC + 1) + 3) / 4 * 4 + 2 * (size_t)kOwnStages * ((CC * PP + 3) / 4 * 4) + 256) * 4;
}
template <int CC>
int launch_owner(const float* top, const float* rois, const int* argmax, float* bottom, int B, int C, int H, int W, int N,
                 int PP, cudaStream_t st) {
  const size_t smem = owner_smem<CC>(H * W, PP);
  // bulk copies need 16-byte aligned, 16-byte multiple tiles: true for whole chunks of CC channels (C % CC == 0)
  const int bulk = (C % CC == 0) && ((CC * PP * 4) % 16 == 0) && aligned16(top) && aligned16(argmax) &&
                   (((size_t)C * PP * 4) % 16 == 0);
  L2S_CUDA_OK(cudaFuncSetAttribute(roi_maxpool_bwd_owner_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_maxpool_bwd_owner_kernel<CC><<<dim3((C + CC - 1) / CC, B), 256, smem, st>>>(top, rois, argmax, bottom, B, C, H, W, N, PP,
                                                                                 bulk);
  L2S_LAUNCH_OK("roi_maxpool_bwd_owner_kernel");
  count_launch();
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" int l2s_roi_maxpool_fwd(int ph, int pw, float scale, const float* features, const float* rois,
                                   float* output, int32_t* argmax, int B, int C, int H, int W, int N,
                                   l2s_stream_t stream) {
  L2S_REQUIRE(features && rois && output, L2S_ERR_ARG, "roi_maxpool_fwd: null pointer");
  L2S_REQUIRE(ph > 0 && pw > 0 && B > 0 && C > 0 && H > 0 && W > 0 && N >= 0, L2S_ERR_SHAPE,
              "roi_maxpool_fwd: bad shape");
  L2S_REQUIRE((size_t)C * H * W < 2147483647u, L2S_ERR_SHAPE, "roi_maxpool_fwd: C*H*W overflows the int32 argmax");
  if (N == 0) return L2S_OK;
  const size_t total = (size_t)N * C * ph * pw;
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count() * 32);
  roi_maxpool_fwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(features, rois, output, argmax, B, C, H, W,
                                                                       N, ph, pw, scale);
  L2S_LAUNCH_OK("roi_maxpool_fwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_roi_maxpool_bwd(int ph, int pw, float scale, const float* top_grad, const float* rois,
                                   float* bottom_grad, const int32_t* argmax, int B, int C, int H, int W, int N,
                                   l2s_stream_t stream) {
  (void)scale;
  L2S_REQUIRE(rois && bottom_grad && (N == 0 || (top_grad && argmax)), L2S_ERR_ARG, "roi_maxpool_bwd: null pointer");
  L2S_REQUIRE(ph > 0 && pw > 0 && B > 0 && C > 0 && H > 0 && W > 0 && N >= 0, L2S_ERR_SHAPE,
              "roi_maxpool_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  // L2S_ROIPOOL_BWD_DETERMINISTIC=1: the owner-computes kernel (bit-reproducible, 8.8 ms at cfg-2 sizes) instead of the
  // atomic scatter (2.6 ms, sums in arrival order: equal within 1e-6 relative run to run)
  const bool owner = env_flag("L2S_ROIPOOL_BWD_DETERMINISTIC");   // read per call: the tests toggle it
  if (N > 0 && owner) {
    const size_t cap = (size_t)max_smem_optin();
    const int HW = H * W, PP = ph * pw;
    if (owner_smem<32>(HW, PP) <= cap) return launch_owner<32>(top_grad, rois, argmax, bottom_grad, B, C, H, W, N, PP, st);
    if (owner_smem<16>(HW, PP) <= cap) return launch_owner<16>(top_grad, rois, argmax, bottom_grad, B, C, H, W, N, PP, st);
    if (owner_smem<8>(HW, PP) <= cap) return launch_owner<8>(top_grad, rois, argmax, bottom_grad, B, C, H, W, N, PP, st);
  }
  L2S_CUDA_OK(cudaMemsetAsync(bottom_grad, 0, (size_t)B * C * H * W * sizeof(float), st));
  if (N == 0) return L2S_OK;
  const size_t total = (size_t)N * C * ph * pw;
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count() * 32);
  roi_maxpool_bwd_kernel<<<blocks, threads, 0, st>>>(top_grad, rois, argmax, bottom_grad, B, C, H, W, N, ph, pw);
  L2S_LAUNCH_OK("roi_maxpool_bwd_kernel");
  count_launch();
  return L2S_OK;
}

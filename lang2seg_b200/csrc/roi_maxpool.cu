// Caffe-style RoI max-pool forward / backward for sm_100a (the reference's POOLING_MODE != 'crop').
//
// Semantics: pyutils/mask-faster-rcnn/lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.cu
//   forward :15-75  -- round(roi*scale), float bin sizes, [floor, ceil) bins clipped to the map,
//                      first strict maximum in row-major order, argmax = flat (c*H+h)*W+w, empty
//                      bin -> 0 / -1.  Integer contract: bins and argmax are bit-exact.
//   backward:104-179 -- the reference gathers over all ROIs per input element (O(R) each); the
//                      same sums are produced here by scattering top_diff to argmax.
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

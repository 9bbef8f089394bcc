// Box head glue of Network._region_classification (network_cycle_response.py:277-290): the 7x7 spatial mean in front
// of the two Linears and the softmax / argmax behind them.  The Linears themselves run stacked as ONE tcgen05 GEMM
// (functional.linear) from the Python side; what is left is pure bandwidth:
//   spatial_mean  (N,C,P) -> (N,C): a warp takes 32 consecutive (n,c) rows = one contiguous 32*P-float block, staged
//                 through shared memory so that global traffic is fully coalesced (a lane-per-row read would touch
//                 32 sectors per instruction); the reference takes mean(3) then mean(2), reproduced for P = 49.
//   softmax_argmax  rows of ncls scores -> probabilities + first maximum (torch.max(...,1)[1]).
#include "common.cuh"

namespace l2s {
namespace {

constexpr int kMeanWarps = 8;

__global__ void __launch_bounds__(kMeanWarps * 32)
spatial_mean_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int P, int side) {
  extern __shared__ float sm[];                      // [kMeanWarps][32 * P]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* tile = sm + (size_t)wid * 32 * P;
  for (int64_t r0 = ((int64_t)blockIdx.x * kMeanWarps + wid) * 32; r0 < rows; r0 += (int64_t)gridDim.x * kMeanWarps * 32) {
    const int nr = (int)min((int64_t)32, rows - r0);
    const float* src = x + r0 * P;
    for (int i = lane; i < nr * P; i += 32) tile[i] = __ldg(src + i);
    __syncwarp();
    if (lane < nr) {
      const float* row = tile + lane * P;
      float tot = 0.f;
      if (side > 0) {                                // mean over W, then over H (network_cycle_response.py:278)
        for (int y = 0; y < side; ++y) {
          float rs = 0.f;
          for (int xx = 0; xx < side; ++xx) rs += row[y * side + xx];
          tot += rs / (float)side;
        }
        tot /= (float)side;
      } else {
        for (int i = 0; i < P; ++i) tot += row[i];
        tot /= (float)P;
      }
      out[r0 + lane] = tot;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kMeanWarps * 32)
spatial_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int64_t rows, int P) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float inv = 1.f / (float)P;
  for (int64_t r0 = ((int64_t)blockIdx.x * kMeanWarps + wid) * 32; r0 < rows; r0 += (int64_t)gridDim.x * kMeanWarps * 32) {
    const int nr = (int)min((int64_t)32, rows - r0);
    const float g = lane < nr ? __ldg(dout + r0 + lane) * inv : 0.f;
    float* dst = dx + r0 * P;
    const int tot = nr * P;
    for (int base = 0; base < tot; base += 32) {          // uniform trip count: the shuffle needs every lane
      const int i = base + lane;
      const float v = __shfl_sync(0xffffffffu, g, min(i, tot - 1) / P);
      if (i < tot) dst[i] = v;
    }
  }
}

// one warp per row
__global__ void __launch_bounds__(256)
softmax_argmax_kernel(const float* __restrict__ score, int64_t ld, float* __restrict__ prob, int64_t* __restrict__ pred,
                      int R, int ncls) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = score + (size_t)r * ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = lane; c < ncls; c += 32) {
    const float v = __ldg(row + c);
    if (v > m) { m = v; mi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float l = 0.f;
  for (int c = lane; c < ncls; c += 32) l += expf(__ldg(row + c) - m);
  l = warp_sum(l);
  const float inv = 1.f / l;
  if (prob)
    for (int c = lane; c < ncls; c += 32) prob[(size_t)r * ncls + c] = expf(__ldg(row + c) - m) * inv;
  if (pred && lane == 0) pred[r] = mi;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" int l2s_spatial_mean_fwd(const float* x, float* out, int64_t rows, int P, l2s_stream_t stream) {
  L2S_REQUIRE(rows >= 0 && P > 0 && P <= 256, L2S_ERR_SHAPE, "spatial_mean: bad shape rows=%lld P=%d", (long long)rows, P);
  if (rows == 0) return L2S_OK;
  L2S_REQUIRE(x && out, L2S_ERR_ARG, "spatial_mean: null pointer");
  int side = 0;
  for (int s = 1; s * s <= P; ++s)
    if (s * s == P) side = s;
  const size_t smem = (size_t)kMeanWarps * 32 * P * sizeof(float);
  L2S_CUDA_OK(cudaFuncSetAttribute(spatial_mean_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t groups = (rows + 32 * kMeanWarps - 1) / (32 * kMeanWarps);
  const int blocks = (int)(groups < (int64_t)sm_count() * 8 ? groups : (int64_t)sm_count() * 8);
  spatial_mean_fwd_kernel<<<blocks, kMeanWarps * 32, smem, (cudaStream_t)stream>>>(x, out, rows, P, side);
  L2S_LAUNCH_OK("spatial_mean_fwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_spatial_mean_bwd(const float* dout, float* dx, int64_t rows, int P, l2s_stream_t stream) {
  L2S_REQUIRE(rows >= 0 && P > 0 && P <= 256, L2S_ERR_SHAPE, "spatial_mean_bwd: bad shape");
  if (rows == 0) return L2S_OK;
  L2S_REQUIRE(dout && dx, L2S_ERR_ARG, "spatial_mean_bwd: null pointer");
  const int64_t groups = (rows + 32 * kMeanWarps - 1) / (32 * kMeanWarps);
  const int blocks = (int)(groups < (int64_t)sm_count() * 8 ? groups : (int64_t)sm_count() * 8);
  spatial_mean_bwd_kernel<<<blocks, kMeanWarps * 32, 0, (cudaStream_t)stream>>>(dout, dx, rows, P);
  L2S_LAUNCH_OK("spatial_mean_bwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_softmax_argmax(const float* score, int64_t ld, float* prob, int64_t* pred, int R, int ncls,
                                  l2s_stream_t stream) {
  L2S_REQUIRE(R >= 0 && ncls > 0 && ld >= ncls, L2S_ERR_SHAPE, "softmax_argmax: bad shape R=%d ncls=%d", R, ncls);
  if (R == 0) return L2S_OK;
  L2S_REQUIRE(score && (prob || pred), L2S_ERR_ARG, "softmax_argmax: null pointer");
  softmax_argmax_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(score, ld, prob, pred, R, ncls);
  L2S_LAUNCH_OK("softmax_argmax_kernel");
  count_launch();
  return L2S_OK;
}

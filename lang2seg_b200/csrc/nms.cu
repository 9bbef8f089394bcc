// Greedy NMS entirely on the device, for sm_100a.
//
// Semantics: gpu_nms of the reference (pyutils/mask-faster-rcnn/lib/nms/src/nms_cuda.c:17-67 with the bit-mask kernel
// nms/src/cuda/nms_kernel.cu:15-83): boxes (N,5) [x1,y1,x2,y2,score] already sorted by descending score; box i
// suppresses every later box j with IoU(i,j) > thresh, IoU with the "+1" pixel convention of devIoU (:15-24).
// The reference computes the N x N/64 bit mask on the GPU, copies it to the host (18 MB at N = 12000) and runs the
// greedy scan on the CPU.  Here the scan stays on the device, so the proposal layer needs no host round trip:
//   nms_mask_kernel    upper-triangular 64 x 64 tiles only (the scan never reads the lower triangle)
//   nms_scan_kernel    one CTA.  Per 64-box chunk: one thread resolves the chunk against its diagonal tile (kept
//                      bits K), then every thread ORs the mask rows of the kept boxes into its words of the
//                      removed-set (up to 64 independent coalesced loads in flight per thread); the next chunk's
//                      diagonal tile is prefetched meanwhile.  keep[] is written in order, num_out at the end.
// IoU uses explicitly rounded fp32 operations (no FMA contraction) so that it is bit-identical to the IEEE fp32
// restatement in oracle/restate.py.
#include "common.cuh"

namespace l2s {
namespace {

constexpr int kTile = 64;

__device__ __forceinline__ float iou_plus1(const float4& a, const float4& b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  const float sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
  const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float* __restrict__ boxes, int n, float thresh, unsigned long long* __restrict__ mask, int col_blocks) {
  // blockIdx.x enumerates the upper-triangular tile pairs (row <= col)
  int rb = 0, rem = blockIdx.x;
  while (rem >= col_blocks - rb) { rem -= col_blocks - rb; ++rb; }
  const int cb = rb + rem;
  __shared__ float4 cbox[kTile];
  const int t = threadIdx.x;
  const int cj = cb * kTile + t;
  if (cj < n) {
    const float* p = boxes + (size_t)cj * 5;
    cbox[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
  }
  __syncthreads();
  const int ri = rb * kTile + t;
  if (ri >= n) return;
  const float* p = boxes + (size_t)ri * 5;
  const float4 me = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
  const int ncol = min(kTile, n - cb * kTile);
  unsigned long long bits = 0;
  for (int j = (rb == cb) ? t + 1 : 0; j < ncol; ++j)
    if (iou_plus1(me, cbox[j]) > thresh) bits |= 1ull << j;
  mask[(size_t)ri * col_blocks + cb] = bits;
}

__global__ void __launch_bounds__(1024, 1)
nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int col_blocks, int max_out,
                int64_t* __restrict__ keep, int64_t* __restrict__ num_out) {
  extern __shared__ unsigned long long remv[];          // [col_blocks]
  __shared__ unsigned long long diag[2][kTile];
  __shared__ unsigned long long s_kept;
  __shared__ int s_base;
  const int t = threadIdx.x;
  for (int j = t; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  if (t == 0) s_base = 0;
  auto load_diag = [&](int c, int buf) {
    if (t < kTile) {
      const int r = c * kTile + t;
      diag[buf][t] = (r < n) ? mask[(size_t)r * col_blocks + c] : 0ull;
    }
  };
  load_diag(0, 0);
  __syncthreads();
  for (int c = 0; c < col_blocks; ++c) {
    const int buf = c & 1;
    if (c + 1 < col_blocks) load_diag(c + 1, buf ^ 1);            // in flight during the scan of chunk c
    if (t == 0) {
      unsigned long long word = remv[c], kept = 0ull;
      const int cnt = min(kTile, n - c * kTile);
      for (int i = 0; i < cnt; ++i)
        if (!((word >> i) & 1ull)) {
          kept |= 1ull << i;
          word |= diag[buf][i];
        }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    const int base = s_base;
    if (t < kTile && ((kept >> t) & 1ull)) {
      const int pos = base + __popcll(kept & ((1ull << t) - 1ull));
      if (max_out <= 0 || pos < max_out) keep[pos] = (int64_t)c * kTile + t;
    }
    // OR the rows of the kept boxes into the removed-set of the later chunks
    for (int j = c + 1 + t; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = 0ull, k = kept;
      while (k) {
        const int i = __ffsll((long long)k) - 1;
        k &= k - 1;
        acc |= mask[(size_t)(c * kTile + i) * col_blocks + j];
      }
      remv[j] |= acc;
    }
    __syncthreads();
    if (t == 0) s_base = base + __popcll(kept);
    if (max_out > 0 && base + __popcll(kept) >= max_out) break;     // uniform: every thread sees the same values
    __syncthreads();
  }
  __syncthreads();
  if (t == 0) {
    const int total = s_base;
    *num_out = (int64_t)((max_out > 0 && total > max_out) ? max_out : total);
  }
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_nms_workspace_bytes(int n) {
  const size_t cb = (size_t)((n > 0 ? n : 0) + kTile - 1) / kTile;
  return (size_t)(n > 0 ? n : 0) * cb * sizeof(unsigned long long) + 256;
}

extern "C" int l2s_nms(const float* boxes_sorted, int n, float thresh, int max_out, int64_t* keep, int64_t* num_out,
                       void* workspace, size_t workspace_bytes, l2s_stream_t stream) {
  L2S_REQUIRE(n >= 0, L2S_ERR_SHAPE, "nms: bad box count %d", n);
  L2S_REQUIRE(num_out, L2S_ERR_ARG, "nms: null num_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    L2S_CUDA_OK(cudaMemsetAsync(num_out, 0, sizeof(int64_t), st));
    return L2S_OK;
  }
  L2S_REQUIRE(boxes_sorted && keep, L2S_ERR_ARG, "nms: null pointer");
  L2S_REQUIRE(workspace && workspace_bytes >= l2s_nms_workspace_bytes(n) && aligned16(workspace), L2S_ERR_WORKSPACE,
              "nms: workspace missing, misaligned or too small");
  const int col_blocks = (n + kTile - 1) / kTile;
  const size_t smem = (size_t)col_blocks * sizeof(unsigned long long);
  L2S_REQUIRE(smem <= (size_t)max_smem_optin() - 4096, L2S_ERR_SHAPE, "nms: %d boxes are too many for one scan CTA", n);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(workspace);
  const long long tiles = (long long)col_blocks * (col_blocks + 1) / 2;
  L2S_REQUIRE(tiles <= 0x7fffffffLL, L2S_ERR_SHAPE, "nms: too many boxes");
  nms_mask_kernel<<<(unsigned)tiles, kTile, 0, st>>>(boxes_sorted, n, thresh, mask, col_blocks);
  L2S_LAUNCH_OK("nms_mask_kernel");
  L2S_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<1, 1024, smem, st>>>(mask, n, col_blocks, max_out, keep, num_out);
  L2S_LAUNCH_OK("nms_scan_kernel");
  count_launch(2);
  return L2S_OK;
}

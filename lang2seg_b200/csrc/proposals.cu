// RPN tail and proposal-target arithmetic on the device (reference: MFR/layer_utils/proposal_layer.py:19-68,
// MFR/layer_utils/proposal_target_layer.py:120-205, MFR/model/bbox_transform.py, MFR/utils/bbox.pyx).
//
// The reference mixes GPU tensor ops with host round trips (numpy.random.choice on the host, per-ROI imresize loop, NMS
// mask copied back).  Here the arithmetic that is specific to this path is three small kernels; ordering / selection
// (sort, top-k) stays on the framework's device primitives and nothing goes through the host:
//   proposal_decode   anchors + deltas -> clipped boxes packed with their score as (N,5) rows, ready for l2s_nms
//   roi_gt_overlaps   IoU of every ROI against every ground-truth box -> max overlap, first arg-max
//   bbox_targets      (ex_roi, assigned gt, label) -> normalised (dx,dy,dw,dh) scattered into the label's 4-of-4K slot
// fp32 with the reference's operation order and no FMA contraction (compiled per expression with __f*_rn), so that a
// numpy restatement of the same formulas reproduces the boxes to the last bit except for exp/log (<= 2 ulp).
#include "common.cuh"

namespace l2s {
namespace {

// bbox_transform_inv (bbox_transform.py:38-64) + clip_boxes (:67-83); out row = [x1,y1,x2,y2,score]
__global__ void proposal_decode_kernel(const float* __restrict__ anchors, const float* __restrict__ deltas,
                                       const float* __restrict__ scores, int64_t score_stride, float* __restrict__ out,
                                       int N, float im_h, float im_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(anchors) + i);
  const float4 d = __ldg(reinterpret_cast<const float4*>(deltas) + i);
  const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
  const float pw = __fmul_rn(expf(d.z), w), ph = __fmul_rn(expf(d.w), h);
  const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
  const float xmax = __fsub_rn(im_w, 1.0f), ymax = __fsub_rn(im_h, 1.0f);
  float* o = out + (size_t)i * 5;
  o[0] = fminf(fmaxf(__fsub_rn(pcx, hw), 0.f), xmax);
  o[1] = fminf(fmaxf(__fsub_rn(pcy, hh), 0.f), ymax);
  o[2] = fminf(fmaxf(__fadd_rn(pcx, hw), 0.f), xmax);
  o[3] = fminf(fmaxf(__fadd_rn(pcy, hh), 0.f), ymax);
  o[4] = __ldg(scores + (size_t)i * score_stride);
}

// bbox_overlaps (utils/bbox.pyx): boxes with the +1 pixel convention; thread per ROI, gt boxes staged in shared memory
__global__ void roi_gt_overlaps_kernel(const float* __restrict__ rois, int roi_stride, int roi_off,
                                       const float* __restrict__ gt, int gt_stride, float* __restrict__ max_ov,
                                       int64_t* __restrict__ arg, int N, int G) {
  extern __shared__ float s_gt[];     // [G][4] + area
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const float x1 = __ldg(gt + (size_t)i * gt_stride), y1 = __ldg(gt + (size_t)i * gt_stride + 1);
    const float x2 = __ldg(gt + (size_t)i * gt_stride + 2), y2 = __ldg(gt + (size_t)i * gt_stride + 3);
    s_gt[i * 5 + 0] = x1; s_gt[i * 5 + 1] = y1; s_gt[i * 5 + 2] = x2; s_gt[i * 5 + 3] = y2;
    s_gt[i * 5 + 4] = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* r = rois + (size_t)n * roi_stride + roi_off;
  const float x1 = __ldg(r), y1 = __ldg(r + 1), x2 = __ldg(r + 2), y2 = __ldg(r + 3);
  const float area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
  float best = -1.f;
  int bi = 0;
  for (int k = 0; k < G; ++k) {
    const float iw = __fadd_rn(__fsub_rn(fminf(x2, s_gt[k * 5 + 2]), fmaxf(x1, s_gt[k * 5 + 0])), 1.f);
    float ov = 0.f;
    if (iw > 0.f) {
      const float ih = __fadd_rn(__fsub_rn(fminf(y2, s_gt[k * 5 + 3]), fmaxf(y1, s_gt[k * 5 + 1])), 1.f);
      if (ih > 0.f) {
        const float inter = __fmul_rn(iw, ih);
        const float ua = __fsub_rn(__fadd_rn(area, s_gt[k * 5 + 4]), inter);
        ov = __fdiv_rn(inter, ua);
      }
    }
    if (ov > best) { best = ov; bi = k; }      // first maximum, as numpy / torch.max on the CPU
  }
  max_ov[n] = G > 0 ? best : 0.f;
  arg[n] = bi;
}

// _compute_targets + _get_bbox_regression_labels (proposal_target_layer.py:83-121): targets / weights are (N, 4K), zeroed
// here, the 4 values of row n land in columns 4*label .. 4*label+3 when label > 0
__global__ void bbox_targets_kernel(const float* __restrict__ rois, int roi_stride, int roi_off,
                                    const float* __restrict__ gt, int gt_stride, const int64_t* __restrict__ assign,
                                    const float* __restrict__ labels, float* __restrict__ targets,
                                    float* __restrict__ inside_w, int N, int K, float4 mean, float4 stdv, float4 inw) {
  const int n = blockIdx.x;
  float* trow = targets + (size_t)n * 4 * K;
  float* wrow = inside_w + (size_t)n * 4 * K;
  for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) {
    trow[i] = 0.f;
    wrow[i] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int cls = (int)__ldg(labels + n);
  if (cls <= 0 || cls >= K) return;
  const float* r = rois + (size_t)n * roi_stride + roi_off;
  const float* q = gt + (size_t)__ldg(assign + n) * gt_stride;
  const float ew = __fadd_rn(__fsub_rn(r[2], r[0]), 1.f), eh = __fadd_rn(__fsub_rn(r[3], r[1]), 1.f);
  const float ecx = __fadd_rn(r[0], __fmul_rn(0.5f, ew)), ecy = __fadd_rn(r[1], __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(q[2], q[0]), 1.f), gh = __fadd_rn(__fsub_rn(q[3], q[1]), 1.f);
  const float gcx = __fadd_rn(q[0], __fmul_rn(0.5f, gw)), gcy = __fadd_rn(q[1], __fmul_rn(0.5f, gh));
  const float dx = __fdiv_rn(__fsub_rn(gcx, ecx), ew), dy = __fdiv_rn(__fsub_rn(gcy, ecy), eh);
  const float dw = logf(__fdiv_rn(gw, ew)), dh = logf(__fdiv_rn(gh, eh));
  trow[4 * cls + 0] = __fdiv_rn(__fsub_rn(dx, mean.x), stdv.x);
  trow[4 * cls + 1] = __fdiv_rn(__fsub_rn(dy, mean.y), stdv.y);
  trow[4 * cls + 2] = __fdiv_rn(__fsub_rn(dw, mean.z), stdv.z);
  trow[4 * cls + 3] = __fdiv_rn(__fsub_rn(dh, mean.w), stdv.w);
  wrow[4 * cls + 0] = inw.x; wrow[4 * cls + 1] = inw.y; wrow[4 * cls + 2] = inw.z; wrow[4 * cls + 3] = inw.w;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" int l2s_proposal_decode(const float* anchors, const float* deltas, const float* scores, int64_t score_stride,
                                   float* boxes5, int N, float im_h, float im_w, l2s_stream_t stream) {
  L2S_REQUIRE(N >= 0 && score_stride >= 1, L2S_ERR_SHAPE, "proposal_decode: bad shape N=%d", N);
  if (N == 0) return L2S_OK;
  L2S_REQUIRE(anchors && deltas && scores && boxes5, L2S_ERR_ARG, "proposal_decode: null pointer");
  L2S_REQUIRE(aligned16(anchors) && aligned16(deltas), L2S_ERR_ALIGN, "proposal_decode: anchors / deltas must be 16-byte aligned");
  proposal_decode_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(anchors, deltas, scores, score_stride, boxes5, N,
                                                                             im_h, im_w);
  L2S_LAUNCH_OK("proposal_decode_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_roi_gt_overlaps(const float* rois, int roi_stride, int roi_offset, const float* gt_boxes, int gt_stride,
                                   float* max_overlap, int64_t* gt_assignment, int N, int G, l2s_stream_t stream) {
  L2S_REQUIRE(N >= 0 && G >= 0 && roi_stride >= roi_offset + 4 && gt_stride >= 4, L2S_ERR_SHAPE, "roi_gt_overlaps: bad shape");
  if (N == 0) return L2S_OK;
  L2S_REQUIRE(rois && max_overlap && gt_assignment && (G == 0 || gt_boxes), L2S_ERR_ARG, "roi_gt_overlaps: null pointer");
  const size_t smem = (size_t)(G > 0 ? G : 1) * 5 * sizeof(float);
  L2S_REQUIRE(smem <= 48 * 1024, L2S_ERR_SHAPE, "roi_gt_overlaps: too many ground-truth boxes (%d)", G);
  roi_gt_overlaps_kernel<<<(N + 127) / 128, 128, smem, (cudaStream_t)stream>>>(rois, roi_stride, roi_offset, gt_boxes,
                                                                               gt_stride, max_overlap, gt_assignment, N, G);
  L2S_LAUNCH_OK("roi_gt_overlaps_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_bbox_targets(const float* rois, int roi_stride, int roi_offset, const float* gt_boxes, int gt_stride,
                                const int64_t* gt_assignment, const float* labels, float* targets, float* inside_weights,
                                int N, int num_classes, const float* means4_host, const float* stds4_host,
                                const float* inside4_host, l2s_stream_t stream) {
  L2S_REQUIRE(N >= 0 && num_classes > 0 && roi_stride >= roi_offset + 4 && gt_stride >= 4, L2S_ERR_SHAPE, "bbox_targets: bad shape");
  if (N == 0) return L2S_OK;
  L2S_REQUIRE(rois && gt_boxes && gt_assignment && labels && targets && inside_weights && means4_host && stds4_host &&
                  inside4_host, L2S_ERR_ARG, "bbox_targets: null pointer");
  const float4 m = make_float4(means4_host[0], means4_host[1], means4_host[2], means4_host[3]);
  const float4 s = make_float4(stds4_host[0], stds4_host[1], stds4_host[2], stds4_host[3]);
  const float4 w = make_float4(inside4_host[0], inside4_host[1], inside4_host[2], inside4_host[3]);
  bbox_targets_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(rois, roi_stride, roi_offset, gt_boxes, gt_stride, gt_assignment,
                                                          labels, targets, inside_weights, N, num_classes, m, s, w);
  L2S_LAUNCH_OK("bbox_targets_kernel");
  count_launch();
  return L2S_OK;
}

// Target construction on the device: crop a ground-truth mask to a box and resize it with PIL-NEAREST semantics.
//
// Replaces the per-ROI host loop of the proposal target layer
// (pyutils/mask-faster-rcnn/lib/layer_utils/proposal_target_layer.py:193-201:
//    cropped = gt_masks[gt_assignment[i], int(y1):int(y2)+1, int(x1):int(x2)+1]
//    cropped = imresize(cropped, (MASK_SIZE, MASK_SIZE), interp='nearest').astype(float32) )
// and, with no boxes, the response target of network_cycle_response.py:418
//    (imresize(gt_mask, response.size(), interp='nearest')),
// both of which run scipy on the CPU and copy the result to the GPU every step.
// Integer contract (bit exact): destination index i of a length-`dst` axis reads source index
// floor((i + 0.5) * src / dst) = ((2i + 1) * src) / (2 * dst) in integer arithmetic.
#include "common.cuh"

namespace l2s {
namespace {

__global__ void __launch_bounds__(256)
mask_crop_resize_kernel(const uint8_t* __restrict__ masks, const float* __restrict__ rois, int roi_stride,
                        const int32_t* __restrict__ assign, float* __restrict__ out, int G, int imH, int imW, int n,
                        int outH, int outW) {
  const int i = blockIdx.x;
  int g = assign ? __ldg(assign + i) : i;
  const bool ok = g >= 0 && g < G;
  int x1 = 0, y1 = 0, x2 = imW, y2 = imH;          // [x1, x2) x [y1, y2): the python slice
  if (rois) {
    const float* r = rois + (size_t)i * roi_stride;
    // int() truncates toward zero; python slicing clips the stop at the array size (boxes are inside the image)
    x1 = max((int)__ldg(r + 1), 0);
    y1 = max((int)__ldg(r + 2), 0);
    x2 = min((int)__ldg(r + 3) + 1, imW);
    y2 = min((int)__ldg(r + 4) + 1, imH);
  }
  const int sw = x2 - x1, sh = y2 - y1;
  const uint8_t* m = masks + (size_t)(ok ? g : 0) * imH * imW;
  for (int p = threadIdx.x; p < outH * outW; p += blockDim.x) {
    const int y = p / outW, x = p - y * outW;
    float v = 0.f;
    if (ok && sw > 0 && sh > 0) {
      const int sy = min((int)(((long long)(2 * y + 1) * sh) / (2 * outH)), sh - 1);
      const int sx = min((int)(((long long)(2 * x + 1) * sw) / (2 * outW)), sw - 1);
      v = (float)__ldg(m + (size_t)(y1 + sy) * imW + x1 + sx);
    }
    out[(size_t)i * outH * outW + p] = v;
  }
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" int l2s_mask_crop_resize(const uint8_t* masks, const float* rois, int roi_stride, const int32_t* assign,
                                    float* out, int G, int imH, int imW, int n, int outH, int outW,
                                    l2s_stream_t stream) {
  L2S_REQUIRE(n >= 0 && G > 0 && imH > 0 && imW > 0 && outH > 0 && outW > 0, L2S_ERR_SHAPE,
              "mask_crop_resize: bad shape n=%d G=%d image %dx%d out %dx%d", n, G, imH, imW, outH, outW);
  if (n == 0) return L2S_OK;
  L2S_REQUIRE(masks && out, L2S_ERR_ARG, "mask_crop_resize: null pointer");
  L2S_REQUIRE(!rois || roi_stride >= 5, L2S_ERR_ARG, "mask_crop_resize: rois rows are [batch,x1,y1,x2,y2,...]");
  mask_crop_resize_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(masks, rois, roi_stride, assign, out, G, imH, imW, n, outH,
                                                               outW);
  L2S_LAUNCH_OK("mask_crop_resize_kernel");
  count_launch();
  return L2S_OK;
}

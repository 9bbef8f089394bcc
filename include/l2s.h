/* l2s.h -- C ABI of libl2s.so: the B200 (sm_100a) kernels behind lang2seg's
 * language-conditioned segmentation hot path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are dense row-major ("NCHW" where a shape is given), fp32 unless noted;
 *   - the caller owns every buffer (outputs, saved state, workspace); the library never
 *     allocates, frees or retains pointers across calls;
 *   - calls are asynchronous on `stream` (a cudaStream_t), re-entrant across streams and
 *     devices, CUDA-graph capturable (no host sync, no malloc);
 *   - return value: 0 on success, negative L2S_ERR_* otherwise; the message is available
 *     from l2s_last_error_string() (per host thread).  The library never calls exit() and
 *     never throws across this boundary.
 *
 * Each entry point cites the reference interface it replaces; paths are relative to the
 * lang2seg reference tree, MFR = pyutils/mask-faster-rcnn/lib.
 */
#ifndef L2S_H_
#define L2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* l2s_stream_t; /* cudaStream_t */

#define L2S_OK 0
#define L2S_ERR_SHAPE (-1)       /* bad / unsupported size */
#define L2S_ERR_ALIGN (-2)       /* pointer not aligned as required */
#define L2S_ERR_WORKSPACE (-3)   /* workspace missing or too small */
#define L2S_ERR_CUDA (-4)        /* CUDA runtime / launch error */
#define L2S_ERR_ARG (-5)         /* null pointer / invalid flag */

#define L2S_NUM_FILTERS 7

/* flags */
#define L2S_GATE_SIGMOID 0  /* Y = X * sigmoid(r)  network_cycle_response.py:570            */
#define L2S_GATE_LINEAR 1   /* Y = X * r           network_7f.py:534, network_cycle_res5_2.py:562 */
#define L2S_CROP_MAX_POOL 1 /* 2S x 2S samples then 2x2 max (network_cycle_response.py:140-144) */
#define L2S_CROP_ALIGN 2    /* _crop_pool_layer_align (network_cycle_response.py:151-182)  */
#define L2S_CROP_BWD_RANKED 4 /* backward only: force the sample-per-lane (ranked) kernel where the row-owner kernel
                               * (lane = channel, warp = map rows; 7x7 crops of maps <= ~1300 pixels) is the default */
#define L2S_CROP_WS_PREPARED 8 /* backward only: `workspace` is the very buffer l2s_roi_crop_fwd filled for the same rois,
                                * sizes and flags (untouched since): the ROI binning and geometry records in it are
                                * reused instead of recomputed (3 small launches, ~64 us at 12288 ROIs) */

/* precision of the tensor-core GEMMs (mask head, caption projections, l2s_gemm_bf16x3) */
#define L2S_PRECISION_FP32 0 /* bf16x3 split products: ~1e-5 from an fp32 GEMM (the default; the 1e-4 parity contract) */
#define L2S_PRECISION_BF16 1 /* one tensor pass on the bf16-rounded operands: north_star's "bf16 variants within 1e-2" */

int l2s_version(void);
/* Process-wide precision switch (applies to GEMMs launched after the call); returns L2S_ERR_ARG for an unknown mode. */
int l2s_set_precision(int mode);
int l2s_get_precision(void);
const char* l2s_last_error_string(void);
/* number of kernels this library launched since load (process wide) */
uint64_t l2s_launch_count(void);
/* Diagnostics: hand the library a device buffer of >= 8192 bytes (or NULL to stop); CTA 0 of the persistent decode
 * kernels then writes %globaltimer at its phase boundaries (slot 0: forward, slot 1: backward; [64 steps][8 marks] u64
 * each).  Not for production use: the pointer is process-global. */
int l2s_set_debug_buffer(void* device_buffer, size_t bytes);

/* ---------------------------------------------------------------------------------------
 * (1) Spatial dynamic filter response + fusion + gate.
 * Replaces MFR/nets/network_cycle_response.py:534-570 (7 masked 1x1 F.conv2d, cat, fuse conv,
 * sigmoid, gate) and, through resp_target, the response loss at :415-422.
 *
 *   X        (I,C,H,W)   backbone C4 map of each image
 *   filt     (E,7,C)     f_k = tanh(dynamic_fc_k(hidden))   (:510-529)
 *   fuse     (E,7)       w   = tanh(response_fc(hidden))    (:531-532)
 *   expr2img (E) int32   image of each expression, NON-DECREASING (expressions grouped by image)
 *   response (E,H,W)     pre-sigmoid fused response  (_predictions['response'], :568)
 *   rk_saved (E,7,H,W) or NULL: the masked per-filter responses r_k, kept for the backward
 *   Y        (E,C,H,W)   gated features (:570)
 *   resp_target (E,H,W) or NULL ; resp_loss (E) or NULL: per-expression mean BCE-with-logits.
 *   workspace: l2s_dynfilter_fwd_workspace_bytes() (bf16 planes of the stacked filters + loss partials of the
 *   tensor-core kernel); with workspace == NULL the call still works and takes the FFMA kernel.
 *   Kernel: TMA-staged [C x 32 px] feature tiles resident in shared memory, the contraction on tcgen05
 *   (bf16x3 split, fp32 accumulation in TMEM), masks / fusion / sigmoid / BCE in the TMEM epilogue, gating streamed
 *   from the resident tile (csrc/dynfilter_tc.cu); shapes outside its range (H*W % 4, C % 32, C > 1024) run the FFMA
 *   tile kernel of csrc/dynfilter.cu.
 * ------------------------------------------------------------------------------------- */
size_t l2s_dynfilter_fwd_workspace_bytes(int I, int E, int C, int H, int W);
int l2s_dynfilter_fwd(const float* X, const float* filt, const float* fuse, const int32_t* expr2img,
                      float* response, float* rk_saved, float* Y, const float* resp_target,
                      float* resp_loss, int I, int E, int C, int H, int W, int flags, void* workspace,
                      size_t workspace_bytes, l2s_stream_t stream);

/* Backward of the above.
 *   rk_saved: what the forward kept, or NULL (then r_k is recomputed into the workspace) ;
 *   dY (E,C,H,W) ; dresponse (E,H,W) or NULL (explicit upstream gradient on the response) ;
 *   resp_target (E,H,W) + resp_gscale (E) or NULL: adds resp_gscale[e]*(sigmoid(r)-t)/(H*W),
 *   the gradient of the response loss, without materialising it.
 *   dX (I,C,H,W) is OVERWRITTEN with the sum over the image's expressions ;
 *   dfilt (E,7,C), dfuse (E,7) are overwritten.
 *   workspace: l2s_dynfilter_bwd_workspace_bytes(). */
size_t l2s_dynfilter_bwd_workspace_bytes(int I, int E, int C, int H, int W);
int l2s_dynfilter_bwd(const float* X, const float* filt, const float* fuse, const int32_t* expr2img,
                      const float* response, const float* rk_saved, const float* dY,
                      const float* dresponse, const float* resp_target, const float* resp_gscale,
                      float* dX, float* dfilt,
                      float* dfuse, int I, int E, int C, int H, int W, int flags, void* workspace,
                      size_t workspace_bytes, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (2a) Crop-and-resize ROI pooling.  Replaces Network._crop_pool_layer / _crop_pool_layer_align
 * (MFR/nets/network_cycle_response.py:107-182: affine_grid + grid_sample [+ max_pool2d]).
 *
 *   bottom (B,C,H,W) ; rois (N,5) [batch,x1,y1,x2,y2] in image pixels, batch in [0,B)
 *   (the reference always has batch 0; here it selects the expression's map) ;
 *   out (N,C,pool,pool) ; argmax (N,C,pool,pool) uint8 winner of each 2x2 block, only with
 *   L2S_CROP_MAX_POOL (may be NULL when no backward is needed) ;
 *   im_h, im_w: image size, used only with L2S_CROP_ALIGN.
 *   workspace: l2s_roi_crop_workspace_bytes(B,N,flags), 16-byte aligned: ROI binning by batch index and one
 *   geometry record per ROI (corner slots + bilinear fractions + collision ranks), rebuilt by every call.
 * ------------------------------------------------------------------------------------- */
size_t l2s_roi_crop_workspace_bytes(int B, int N, int flags);
int l2s_roi_crop_fwd(const float* bottom, const float* rois, float* out, uint8_t* argmax, int B, int C,
                     int H, int W, int N, int pool, int flags, float im_h, float im_w, void* workspace,
                     size_t workspace_bytes, l2s_stream_t stream);
/* dbottom (B,C,H,W) is OVERWRITTEN with the scatter-add over all ROIs (deterministic order). */
int l2s_roi_crop_bwd(const float* dout, const float* rois, const uint8_t* argmax, float* dbottom, int B,
                     int C, int H, int W, int N, int pool, int flags, float im_h, float im_w,
                     void* workspace, size_t workspace_bytes, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (2b) Caffe RoI max-pool.  Replaces
 *   int roi_pooling_forward_cuda(int pooled_height, int pooled_width, float spatial_scale,
 *        THCudaTensor* features, THCudaTensor* rois, THCudaTensor* output, THCudaIntTensor* argmax)
 *   int roi_pooling_backward_cuda(... top_grad, rois, bottom_grad, argmax)
 * (MFR/layer_utils/roi_pooling/src/roi_pooling_cuda.h:1-4, kernels in
 *  src/cuda/roi_pooling_kernel.cu:15-75,104-179).  Unlike the reference, any batch size B is
 * accepted; argmax is the flat (c*H+h)*W+w index inside the ROI's image, -1 for empty bins.
 * ------------------------------------------------------------------------------------- */
int l2s_roi_maxpool_fwd(int pooled_height, int pooled_width, float spatial_scale, const float* features,
                        const float* rois, float* output, int32_t* argmax, int B, int C, int H, int W,
                        int N, l2s_stream_t stream);
int l2s_roi_maxpool_bwd(int pooled_height, int pooled_width, float spatial_scale, const float* top_grad,
                        const float* rois, float* bottom_grad, const int32_t* argmax, int B, int C,
                        int H, int W, int N, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * GEMM building block of the mask head: fp32-accurate products on the bf16 tensor pipe.
 * Each fp32 operand is split into bf16 (hi, lo) planes; D = Ahi*Bhi + Ahi*Blo + Alo*Bhi is
 * accumulated in fp32 in TMEM by tcgen05.mma (relative error ~2^-16, inside the 1e-4 budget).
 *
 * l2s_split_bf16: src fp32 (rows, cols) with row stride ld_src -> hi, lo bf16 (rows, ld_dst),
 *                 zero-filling columns [cols, ld_dst).
 * l2s_gemm_bf16x3: D (M,N) fp32, ldd = N.
 *     a_layout/b_layout 0: operand stored K-major  (A: [M][K], B: [N][K])
 *                       1: operand stored MN-major (A: [K][M], B: [K][N])
 *     epilogue  0: D = acc ; 1: D += acc ; 2: D = relu(acc + bias[col/bias_div]) ; 4: D = acc + bias[col/bias_div]
 *     split_k >= 1 as given, 0 = the library picks the factor that fills the machine (epilogues 0/1 only);
 *     with split_k != 1 the epilogue accumulates atomically (epilogue 0 zeroes D first).
 *     The CTA shape (128- or 256-row tile, 4 or 8 epilogue warps) is chosen per problem; the environment
 *     variable L2S_GEMM_SHAPE=0|1|2 pins it (parity tests, profiling).
 *     Any M, N, K (tails use TMA out-of-bounds zero fill); operand row strides and pointers must be
 *     16-byte aligned (K % 8 == 0 for K-major operands, rows % 8 == 0 for MN-major ones).
 * ------------------------------------------------------------------------------------- */
int l2s_split_bf16(const float* src, uint16_t* hi, uint16_t* lo, int64_t rows, int64_t cols, int64_t ld_src,
                   int64_t ld_dst, l2s_stream_t stream);
int l2s_gemm_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, const uint16_t* b_hi, const uint16_t* b_lo,
                    float* D, const float* bias, int bias_div, int M, int N, int K, int a_layout,
                    int b_layout, int epilogue, int split_k, l2s_stream_t stream);
/* exact fp32 FFMA GEMM for small / ragged shapes: D[m,n] (+)= sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk] */
int l2s_gemm_f32(const float* A, const float* B, float* D, int M, int N, int K, int64_t sam, int64_t sak,
                 int64_t sbn, int64_t sbk, int64_t ldd, int accumulate, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Mask head.  Replaces Network._mask_prediction (MFR/nets/network_cycle_response.py:292-307:
 * ConvTranspose2d(Cin,Cmid,2,2) -> ReLU -> Conv2d(Cmid,ncls,1) -> sigmoid) and the mask loss
 * (:404-413: gather class channel, BCE-with-logits, mean).
 *
 *   x (n,Cin,7,7) ; up_w (Cin,Cmid,2,2) ; up_b (Cmid) ; pred_w (ncls,Cmid) ; pred_b (ncls)
 *   score, prob (n,ncls,14,14)
 *   saved: caller-owned buffer of l2s_mask_head_saved_bytes() kept until the backward call.
 *   backward: dscore (n,ncls,14,14) -> dx, d_up_w, d_up_b, d_pred_w, d_pred_b (overwritten).
 *   Requirements: Cin % 64 == 0, Cmid % 16 == 0 (reference: 2048, 256, 81).
 * ------------------------------------------------------------------------------------- */
size_t l2s_mask_head_saved_bytes(int n, int Cin, int Cmid, int ncls);
size_t l2s_mask_head_workspace_bytes(int n, int Cin, int Cmid, int ncls);
int l2s_mask_head_fwd(const float* x, const float* up_w, const float* up_b, const float* pred_w,
                      const float* pred_b, float* score, float* prob, void* saved, int n, int Cin,
                      int Cmid, int ncls, void* workspace, size_t workspace_bytes, l2s_stream_t stream);
/* The forward by stage: stages is a subset of 1 (operand repacks -> bf16 planes) | 2 (deconv GEMM + bias + ReLU ->
 * U planes in `saved`) | 4 (1x1 GEMM + bias -> score/prob).  l2s_mask_head_fwd == stages 7; a single stage may be
 * re-run on the buffers a full call left behind (profiling / bench.py's in-step roofline kernel). */
int l2s_mask_head_fwd_stages(const float* x, const float* up_w, const float* up_b, const float* pred_w,
                             const float* pred_b, float* score, float* prob, void* saved, int n, int Cin,
                             int Cmid, int ncls, void* workspace, size_t workspace_bytes, int stages,
                             l2s_stream_t stream);
int l2s_mask_head_bwd(const float* dscore, const float* up_w, const float* pred_w, const void* saved,
                      float* dx, float* d_up_w, float* d_up_b, float* d_pred_w, float* d_pred_b, int n,
                      int Cin, int Cmid, int ncls, void* workspace, size_t workspace_bytes,
                      l2s_stream_t stream);
/* Backward of the mask head when the ONLY gradient on the scores is the mask loss below (the reference's training
 * step: network_cycle_response.py:404-413 reads nothing but channel label_i of mask_score).  dscore is then one
 * non-zero channel per ROI, so dU = g (x) pred_w[label_i] needs no GEMM: one streaming kernel builds the dU planes,
 * d_up_b, d_pred_w and d_pred_b straight from (score, labels, target) -- dscore is never materialised -- and the two
 * big GEMMs (dx, d_up_w) follow as in l2s_mask_head_bwd.  gscale (1): upstream gradient of the scalar loss.
 * Requires Cmid = 8*d with d a divisor of 256 (returns L2S_ERR_SHAPE otherwise: use l2s_mask_bce_bwd +
 * l2s_mask_head_bwd). */
int l2s_mask_head_bce_bwd(const float* score, const int64_t* labels, const float* target, const float* gscale,
                          const float* up_w, const float* pred_w, const void* saved, float* dx, float* d_up_w,
                          float* d_up_b, float* d_pred_w, float* d_pred_b, int n, int Cin, int Cmid, int ncls,
                          void* workspace, size_t workspace_bytes, l2s_stream_t stream);
/* loss = mean_{i,y,x} BCEWithLogits(score[i,label_i,y,x], target[i,y,x]) ; labels int64 (n).
 * bwd writes dscore (n,ncls,hw) = gscale * dloss/dscore (zero outside the label channel). */
int l2s_mask_bce_fwd(const float* score, const int64_t* labels, const float* target, float* loss, int n,
                     int ncls, int hw, l2s_stream_t stream);
int l2s_mask_bce_bwd(const float* score, const int64_t* labels, const float* target, const float* gscale,
                     float* dscore, int n, int ncls, int hw, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * RPN tail / proposal targets.  Replace the arithmetic of
 *   proposal_layer        MFR/layer_utils/proposal_layer.py:41-46 (bbox_transform_inv + clip_boxes, MFR/model/bbox_transform.py)
 *   _sample_rois          MFR/layer_utils/proposal_target_layer.py:137-139 (bbox_overlaps + max), :184-188
 *                         (_compute_targets + _get_bbox_regression_labels)
 * fp32 in the reference's operation order without FMA contraction.
 *   proposal_decode: anchors (N,4), deltas (N,4), scores (N) with element stride score_stride -> boxes5 (N,5)
 *                    [x1,y1,x2,y2,score] clipped to [0, im_w-1] x [0, im_h-1]: the row layout l2s_nms takes.
 *   roi_gt_overlaps: rois (N, roi_stride) with the box at column roi_offset, gt_boxes (G, gt_stride) [x1,y1,x2,y2,..] ->
 *                    max_overlap (N) and gt_assignment (N) int64 = index of the FIRST maximum (+1-pixel IoU convention).
 *   bbox_targets:    targets, inside_weights (N, 4*num_classes) overwritten: zeros except columns 4*label..4*label+3 of
 *                    rows with label > 0 ; labels (N) fp32 ; means/stds/inside weights are HOST arrays of 4 floats.
 * ------------------------------------------------------------------------------------- */
int l2s_proposal_decode(const float* anchors, const float* deltas, const float* scores, int64_t score_stride,
                        float* boxes5, int N, float im_h, float im_w, l2s_stream_t stream);
int l2s_roi_gt_overlaps(const float* rois, int roi_stride, int roi_offset, const float* gt_boxes, int gt_stride,
                        float* max_overlap, int64_t* gt_assignment, int N, int G, l2s_stream_t stream);
int l2s_bbox_targets(const float* rois, int roi_stride, int roi_offset, const float* gt_boxes, int gt_stride,
                     const int64_t* gt_assignment, const float* labels, float* targets, float* inside_weights,
                     int N, int num_classes, const float* means4_host, const float* stds4_host,
                     const float* inside4_host, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Box head glue.  Replaces the pieces of Network._region_classification
 * (MFR/nets/network_cycle_response.py:277-290) around its two Linears: `spatial_fc7.mean(3).mean(2)` and
 * `F.softmax(cls_score, 1)` / `torch.max(cls_score, 1)[1]`.  (The two Linears run stacked as one GEMM on
 * l2s_gemm_bf16x3 / l2s_linear_small.)
 *   x (rows, P) fp32 with rows = N*C and P = 7*7 -> out (rows) ; the square case takes the mean over W then over H
 *   like the reference.  bwd: dx[r, p] = dout[r] / P.
 *   softmax_argmax: score (R, ncls) with row stride ld -> prob (R, ncls) dense and / or pred (R) int64 = index of the
 *   first maximum.  Either output may be NULL.
 * ------------------------------------------------------------------------------------- */
int l2s_spatial_mean_fwd(const float* x, float* out, int64_t rows, int P, l2s_stream_t stream);
int l2s_spatial_mean_bwd(const float* dout, float* dx, int64_t rows, int P, l2s_stream_t stream);
int l2s_softmax_argmax(const float* score, int64_t ld, float* prob, int64_t* pred, int R, int ncls,
                       l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Targets on the device: crop a uint8 {0,1} ground-truth mask to a box and resize it
 * to (outH,outW) with scipy.misc.imresize(..., interp='nearest') semantics.  Replaces the host loop of
 * MFR/layer_utils/proposal_target_layer.py:193-201 (mask targets: masks[assign[i], int(y1):int(y2)+1,
 * int(x1):int(x2)+1] -> 14x14) and, with rois == NULL, the response target of
 * MFR/nets/network_cycle_response.py:418 (whole mask -> response size).
 *   masks (G,imH,imW) uint8 ; rois (n,roi_stride) fp32 rows [batch,x1,y1,x2,y2,...] in image pixels or NULL ;
 *   assign (n) int32 index of each output's mask, or NULL (output i uses mask i) ; out (n,outH,outW) fp32.
 *   Bit exact: destination index i reads source index floor((i+0.5)*src/dst).  Empty crops give zeros.
 * ------------------------------------------------------------------------------------- */
int l2s_mask_crop_resize(const uint8_t* masks, const float* rois, int roi_stride, const int32_t* assign, float* out,
                         int G, int imH, int imW, int n, int outH, int outW, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Greedy NMS on the device.  Replaces
 *   int gpu_nms(THLongTensor* keep, THLongTensor* num_out, THCudaTensor* boxes, float nms_overlap_thresh)
 * (MFR/nms/src/nms_cuda.c:17-67; bit-mask kernel MFR/nms/src/cuda/nms_kernel.cu:26-83, IoU with the +1 pixel
 * convention :15-24), whose greedy scan runs on the HOST after copying the N x N/64 mask back.
 *   boxes_sorted (n,5) [x1,y1,x2,y2,score], sorted by descending score (as pth_nms.py:36 does) ;
 *   keep (n) int64 and num_out (1) int64 are DEVICE buffers (the reference's are host tensors): keep[0..num_out)
 *   are the indices (into boxes_sorted) of the surviving boxes in score order ;
 *   max_out > 0 stops after that many survivors (RPN_POST_NMS_TOP_N), 0 keeps all ;
 *   workspace: l2s_nms_workspace_bytes(n) (the bit mask), 16-byte aligned.
 * ------------------------------------------------------------------------------------- */
size_t l2s_nms_workspace_bytes(int n);
int l2s_nms(const float* boxes_sorted, int n, float thresh, int max_out, int64_t* keep, int64_t* num_out,
            void* workspace, size_t workspace_bytes, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (3) att2in2 attention step.  Replaces Attention.forward
 * (lib/caption_models/AttModel.py:406-423) after the h2att Linear:
 *   e_a = alpha_w . tanh(p_att[b,a,:] + att_h[b,:]) + alpha_b ; weight = softmax_a(e) ;
 *   att_res[b,:] = sum_a weight[b,a] * att_feats[b,a,:]
 *
 *   att_h (B,Dh) ; att_feats (B,A,D) ; p_att (B,A,Dh) ; alpha_w (Dh) ; alpha_b (1)
 *   weight (B,A) and att_res (B,D) are outputs.  D, Dh multiples of 4 and <= 1024.
 *   backward: datt_res (B,D) ->
 *     datt_h (B,Dh) overwritten ; de (B,A) overwritten (gradient on the scores) ;
 *     dp_att (B,A,Dh) ACCUMULATED (+=) when non-NULL ; datt_feats (B,A,D) ACCUMULATED when
 *     non-NULL ; dalpha_w (Dh) ACCUMULATED atomically when non-NULL.
 * ------------------------------------------------------------------------------------- */
int l2s_att_step_fwd(const float* att_h, const float* att_feats, const float* p_att, const float* alpha_w,
                     const float* alpha_b, float* weight, float* att_res, int B, int A, int D, int Dh,
                     l2s_stream_t stream);
int l2s_att_step_bwd(const float* datt_res, const float* att_h, const float* att_feats, const float* p_att,
                     const float* alpha_w, const float* weight, float* datt_h, float* de, float* dp_att,
                     float* datt_feats, float* dalpha_w, int B, int A, int D, int Dh, l2s_stream_t stream);

/* att2in2 gate epilogue.  Replaces Att2in2Core.forward lines AttModel.py:450-464 after the three
 * Linears: sums (B,5D) = i2h(xt)+h2h(h) ; a2c_out (B,2D) = a2c(att_res) ; c_prev (B,D)
 *   i,f,o = sigmoid(sums[:, 0:3D]) ; g = max-halves(sums[:,3D:5D] + a2c_out)
 *   c = f*c_prev + i*g ; h = o*tanh(c)
 * backward: dh, dc (B,D) -> dsums (B,5D), da2c (B,2D), dc_prev (B,D). */
int l2s_att2in2_gates_fwd(const float* sums, const float* a2c_out, const float* c_prev, float* h, float* c,
                          int B, int D, l2s_stream_t stream);
int l2s_att2in2_gates_bwd(const float* sums, const float* a2c_out, const float* c_prev, const float* c,
                          const float* dh, const float* dc, float* dsums, float* da2c, float* dc_prev,
                          int B, int D, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * att2in2 decode loop.  Replaces the T-step recurrence of AttModel.forward
 * (lib/caption_models/AttModel.py:75-99) with Att2in2Core.forward (:446-466) and Attention.forward
 * (:406-423) inside it, as one call per direction (4 kernel launches per token, issued by the library).
 * The caller hoists what does not depend on the state: i2h(xt) for all t, logit/log-softmax/NLL for
 * all t, and every weight gradient (GEMMs over the stacked T*B rows of the buffers below).
 *
 *   LC = Dh + 5*D ; requires D == Dh (both rnn_size = att_hid_size = 512 in att2in2).
 *   cat_all  (T,B,LC)  in/out: on entry row (t,b) = [ b_h2att | i2h(x_t)+b_i2h+b_h2h ] ; on exit
 *                      [ att_h_t | sums_t ] (h_{t-1} . [W_h2att;W_h2h]^T added for t >= 1) -- kept for bwd
 *   w_cat    (LC,D)    rows [h2att.weight ; h2h.weight] ;  w_a2c (2D,D), b_a2c (2D)
 *   alpha_w (Dh), alpha_b (1) ; att_feats (B,A,D), p_att (B,A,Dh)
 *   outputs / saved: h_all, c_all (T,B,D) ; a2c_all (T,B,2D) ; pi_all (T,B,A) softmax weights ;
 *                    att_res_all (T,B,D)
 * backward: dh_all (T,B,D) gradient on every h_t (from the logit layer) ;
 *   w_cat_t (D,LC) and w_a2c_t (D,2D) are the TRANSPOSED weights ;
 *   dcat_all (T,B,LC) = [datt_h_t | dsums_t] ; da2c_all (T,B,2D) ; dres_all (T,B,D) = datt_res_t ;
 *   de_all (T,B,A) score gradients ; dp_att, datt_feats (B,A,D) and dalpha_w (Dh) are OVERWRITTEN.
 *   Weight gradients follow as  dW_cat = dcat_all[1:]^T . h_all[:-1],  dW_a2c = da2c_all^T . att_res_all,
 *   db_h2att = sum dcat[:, :Dh], d(i2h) = dcat[:, Dh:], db_a2c = sum da2c, dalpha_b = sum de_all.
 * workspace: l2s_att2in2_decode_workspace_bytes() for both directions.
 * ------------------------------------------------------------------------------------- */
size_t l2s_att2in2_decode_workspace_bytes(int T, int B, int A, int D, int Dh);
int l2s_att2in2_decode_fwd(float* cat_all, const float* att_feats, const float* p_att, const float* w_cat,
                           const float* w_a2c, const float* b_a2c, const float* alpha_w, const float* alpha_b,
                           float* h_all, float* c_all, float* a2c_all, float* pi_all, float* att_res_all, int T,
                           int B, int A, int D, int Dh, void* workspace, size_t workspace_bytes,
                           l2s_stream_t stream);
int l2s_att2in2_decode_bwd(const float* dh_all, const float* cat_all, const float* att_feats, const float* p_att,
                           const float* w_cat_t, const float* w_a2c_t, const float* alpha_w, const float* c_all,
                           const float* a2c_all, const float* pi_all, float* dcat_all, float* da2c_all,
                           float* dres_all, float* de_all, float* dp_att, float* datt_feats, float* dalpha_w, int T,
                           int B, int A, int D, int Dh, void* workspace, size_t workspace_bytes,
                           l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Variable-length bidirectional LSTM recurrence of the expression encoder.  Replaces
 * pack_padded_sequence -> nn.LSTM(bidirectional, 1 layer) -> pad_packed_sequence in RNNEncoder.forward
 * (lib/layers/lang_encoder.py:38-80) by masking: sequence b advances only while t < lens[b].
 *
 *   G     (B,L,2,4H) in/out: on entry x_t W_ih^T + b_ih + b_hh per direction (gate order i,f,g,o); on exit the
 *                            full pre-activations (h W_hh^T added) -- kept for the backward call
 *   w_hh  (2,4H,H)   weight_hh_l0, weight_hh_l0_reverse ;  lens (B) int32 expression lengths (1..L)
 *   c_all, h_all (L,2,B,H) states after every TIME step (saved) ; out (B,L,2H) zero where t >= lens[b] ;
 *   hidden (B,2H) = [h_fwd(last) | h_bwd(first)]
 * backward: dout (B,L,2H) and/or dhidden (B,2H) -> dG (B,L,2,4H) gradient on the pre-activations (= gradient on
 *   the entry value of G) ; w_hh_t (2,H,4H) TRANSPOSED recurrent weights.
 *   dW_hh[dir] = sum_t dG[:,t,dir]^T . h_prev(t,dir) is left to the caller (one GEMM per direction).
 * ------------------------------------------------------------------------------------- */
size_t l2s_bilstm_workspace_bytes(int L, int B, int H);
int l2s_bilstm_fwd(float* G, const float* w_hh, const int32_t* lens, float* c_all, float* h_all, float* out,
                   float* hidden, int L, int B, int H, void* workspace, size_t workspace_bytes,
                   l2s_stream_t stream);
int l2s_bilstm_bwd(const float* dout, const float* dhidden, const float* G, const float* w_hh_t,
                   const int32_t* lens, const float* c_all, float* dG, int L, int B, int H, void* workspace,
                   size_t workspace_bytes, l2s_stream_t stream);

/* Skinny exact-fp32 linear layer used inside the decode loop (nn.Linear on a batch of <= a few hundred
 * rows): D[M,N] (+)= A[M,K] . W[N,K]^T + bias[N].  K, lda, ldw multiples of 4; deterministic split-K. */
/* out[c] = sum_r in[r*ld + c] (bias gradients of the projections: torch's dim-0 reduction of a tall matrix takes
 * ~20 us per call, this two-stage fixed-order sum ~5 us).  workspace: l2s_colsum_workspace_bytes(R, C). */
size_t l2s_colsum_workspace_bytes(int R, int C);
int l2s_colsum(const float* in, int64_t ld, float* out, int R, int C, void* workspace, size_t workspace_bytes,
               l2s_stream_t stream);
/* Weight gradient of nn.Embedding (AttModel.py:41 `embed`, lang_encoder.py:22 `embedding`): dW[v,:] = sum of dy[r,:] over
 * the rows r with idx[r] == v, in a fixed order (bit-reproducible); rows of dW without a match are zero
 * filled, so the caller does not clear dW.  idx values outside [0,V) are ignored.  D % 4 == 0, 16-byte aligned. */
int l2s_embedding_bwd(const int64_t* idx, const float* dy, float* dW, int R, int V, int D, l2s_stream_t stream);
size_t l2s_linear_small_workspace_bytes(int M, int N, int K);
int l2s_linear_small(const float* A, const float* W, const float* bias, float* D, int M, int N, int K, int lda,
                     int ldw, int ldd, int accumulate, void* workspace, size_t workspace_bytes,
                     l2s_stream_t stream);

/* log-softmax + masked NLL (AttModel.py:98 + lib/misc/utils.py:43-53) on logits (R,V):
 *   logp = log_softmax(logits) (written when non-NULL) ; nll[r] = -logp[r,target[r]]*mask[r]
 * backward: dlogits[r,:] = gscale*mask[r]*(softmax - onehot(target)). */
int l2s_logsoftmax_nll_fwd(const float* logits, const int64_t* target, const float* mask, float* logp,
                           float* nll, int R, int V, l2s_stream_t stream);
int l2s_logsoftmax_nll_bwd(const float* logits, const int64_t* target, const float* mask, const float* gscale,
                           float* dlogits, int R, int V, l2s_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Caption feature prep ("next" row f1).  Replaces network_cycle_response.py:428-432:
 *   fc[b,c] = mean_{hw} feats[b,c,:,:] ; att[b,i,j,c] = adaptive_avg_pool2d(feats,[S,S]) in NHWC,
 * written into column block [col_off, col_off+C) of fc (B,ldc) and att (B,S,S,ldc) so that the
 * before/after concat (:437-438) needs no extra pass.  Backward accumulates into dfeats (=).
 * ------------------------------------------------------------------------------------- */
int l2s_caption_feats_fwd(const float* feats, float* fc, float* att, int B, int C, int H, int W, int S,
                          int ldc, int col_off, l2s_stream_t stream);
int l2s_caption_feats_bwd(const float* dfc, const float* datt, float* dfeats, int B, int C, int H, int W,
                          int S, int ldc, int col_off, l2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* L2S_H_ */

"""torch.autograd bindings of the libl2s.so kernels.

Every function here runs on CUDA tensors through the C ABI of include/l2s.h on the current
torch stream.  There is no CPU path and no PyTorch fallback: a missing library or a CPU tensor
raises.  Shapes are the reference's logical NCHW shapes.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, f32c, ptr, stream

NUM_FILTERS = 7
GATE_SIGMOID, GATE_LINEAR = 0, 1
CROP_MAX_POOL, CROP_ALIGN, CROP_BWD_RANKED, CROP_WS_PREPARED = 1, 2, 4, 8


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# (1) dynamic filter response layer
# ------------------------------------------------------------------------------------------------
class _DynamicFilter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, filt, fuse, expr2img, gate, resp_target):
        X, filt, fuse = f32c(X), f32c(filt), f32c(fuse)
        I, C, H, W = X.shape
        E = filt.shape[0]
        assert filt.shape == (E, NUM_FILTERS, C) and fuse.shape == (E, NUM_FILTERS)
        e2i = expr2img.to(device=X.device, dtype=torch.int32).contiguous()
        response = torch.empty(E, 1, H, W, device=X.device, dtype=torch.float32)
        rk = torch.empty(E, NUM_FILTERS, H, W, device=X.device, dtype=torch.float32)
        Y = torch.empty(E, C, H, W, device=X.device, dtype=torch.float32)
        tgt = f32c(resp_target) if resp_target is not None else None
        loss = torch.empty(E, device=X.device, dtype=torch.float32) if tgt is not None else None
        nbytes = _lib.size("l2s_dynfilter_fwd_workspace_bytes", I, E, C, H, W)
        ws = _ws(nbytes, X.device)
        call("l2s_dynfilter_fwd", ptr(X), ptr(filt), ptr(fuse), ptr(e2i), ptr(response), ptr(rk), ptr(Y),
             ptr(tgt), ptr(loss), I, E, C, H, W, gate, ptr(ws), nbytes, stream())
        ctx.save_for_backward(X, filt, fuse, e2i, response, rk, tgt)
        ctx.gate = gate
        ctx.set_materialize_grads(False)       # unused outputs arrive as None in backward (handled there)
        if loss is None:
            loss = X.new_zeros(E)
        return response, Y, loss

    @staticmethod
    def backward(ctx, dresp, dY, dloss):
        X, filt, fuse, e2i, response, rk, tgt = ctx.saved_tensors
        I, C, H, W = X.shape
        E = filt.shape[0]
        dY = f32c(dY) if dY is not None else torch.zeros(E, C, H, W, device=X.device)
        dresp = f32c(dresp) if dresp is not None else None
        gscale = f32c(dloss) if (dloss is not None and tgt is not None) else None
        dX = torch.empty_like(X)
        dfilt = torch.empty_like(filt)
        dfuse = torch.empty_like(fuse)
        nbytes = _lib.size("l2s_dynfilter_bwd_workspace_bytes", I, E, C, H, W)
        ws = _ws(nbytes, X.device)
        call("l2s_dynfilter_bwd", ptr(X), ptr(filt), ptr(fuse), ptr(e2i), ptr(response), ptr(rk), ptr(dY),
             ptr(dresp), ptr(tgt if gscale is not None else None), ptr(gscale), ptr(dX), ptr(dfilt), ptr(dfuse),
             I, E, C, H, W, ctx.gate, ptr(ws), nbytes, stream())
        return dX, dfilt, dfuse, None, None, None


def dynamic_filter(X, filt, fuse, expr2img=None, gate="sigmoid", resp_target=None):
    """Spatial dynamic filtering + fusion + gate (network_cycle_response.py:534-570).

    X (I,C,H,W) ; filt (E,7,C) ; fuse (E,7) ; expr2img (E,) non-decreasing image index per expression.
    Returns (response (E,1,H,W) pre-sigmoid, Y (E,C,H,W), resp_loss (E,)) where resp_loss is the
    per-expression response BCE (:415-422) against resp_target (E,H,W) (zeros when no target).
    """
    E = filt.shape[0]
    if expr2img is None:
        assert X.shape[0] == E, "expr2img is required when #images != #expressions"
        expr2img = torch.arange(E, device=X.device, dtype=torch.int32)
    elif not (torch.is_tensor(expr2img) and expr2img.is_cuda):
        # a host-side map is validated here for free (the kernels find an image's expressions by its sorted ranges: an
        # unsorted or out-of-range map would silently skip expressions); a device tensor is the caller's contract
        e = torch.as_tensor(expr2img).reshape(-1).long()
        if e.numel() != E or (e.numel() and (int(e.min()) < 0 or int(e.max()) >= X.shape[0] or bool((e[1:] < e[:-1]).any()))):
            raise _lib.L2SError("dynamic_filter: expr2img must hold %d non-decreasing image indices in [0, %d)" % (E, X.shape[0]))
        expr2img = e
    g = GATE_SIGMOID if gate == "sigmoid" else GATE_LINEAR
    return _DynamicFilter.apply(X, filt, fuse, expr2img, g, resp_target)


# ------------------------------------------------------------------------------------------------
# (2a) crop-and-resize ROI pooling
# ------------------------------------------------------------------------------------------------
class _RoICrop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bottom, rois, flags, im_h, im_w, pool):
        bottom = f32c(bottom)
        rois = f32c(rois.detach())
        B, C, H, W = bottom.shape
        N = rois.shape[0]
        assert rois.dim() == 2 and rois.shape[1] == 5, "rois must be (N,5) [batch,x1,y1,x2,y2]"
        out = torch.empty(N, C, pool, pool, device=bottom.device, dtype=torch.float32)
        arg = torch.empty(N, C, pool, pool, device=bottom.device, dtype=torch.uint8) if flags & CROP_MAX_POOL else None
        nbytes = _lib.size("l2s_roi_crop_workspace_bytes", B, N, flags)
        ws = _ws(nbytes, bottom.device)
        call("l2s_roi_crop_fwd", ptr(bottom), ptr(rois), ptr(out), ptr(arg), B, C, H, W, N, pool, flags,
             float(im_h), float(im_w), ptr(ws), nbytes, stream())
        ctx.save_for_backward(rois, arg, ws)      # the workspace keeps the ROI binning + geometry records for the backward
        ctx.meta = (B, C, H, W, N, pool, flags, float(im_h), float(im_w))
        return out

    @staticmethod
    def backward(ctx, dout):
        rois, arg, ws = ctx.saved_tensors
        B, C, H, W, N, pool, flags, im_h, im_w = ctx.meta
        dout = f32c(dout)
        dbottom = torch.empty(B, C, H, W, device=dout.device, dtype=torch.float32)
        call("l2s_roi_crop_bwd", ptr(dout), ptr(rois), ptr(arg), ptr(dbottom), B, C, H, W, N, pool,
             flags | CROP_WS_PREPARED, im_h, im_w, ptr(ws), ws.numel(), stream())
        return dbottom, None, None, None, None, None


def roi_crop(bottom, rois, max_pool=False, align_im_hw=None, pool=7, bwd_ranked=False):
    """Network._crop_pool_layer / _crop_pool_layer_align (network_cycle_response.py:107-182).
    bwd_ranked forces the sample-per-lane backward kernel where the row-owner one (lane = channel) is the default."""
    flags = (CROP_MAX_POOL if max_pool else 0) | (CROP_ALIGN if align_im_hw is not None else 0) | \
            (CROP_BWD_RANKED if bwd_ranked else 0)
    im_h, im_w = align_im_hw if align_im_hw is not None else (0.0, 0.0)
    return _RoICrop.apply(bottom, rois, flags, im_h, im_w, pool)


# ------------------------------------------------------------------------------------------------
# (2b) RoI max-pool
# ------------------------------------------------------------------------------------------------
class _RoIMaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale):
        features = f32c(features)
        rois = f32c(rois.detach())
        B, C, H, W = features.shape
        N = rois.shape[0]
        if rois.dim() != 2 or rois.shape[1] != 5:
            raise _lib.L2SError("roi_pooling: rois must be (N,5)")     # reference returns 0 (roi_pooling_cuda.c:20-23)
        out = torch.empty(N, C, ph, pw, device=features.device, dtype=torch.float32)
        arg = torch.empty(N, C, ph, pw, device=features.device, dtype=torch.int32)
        call("l2s_roi_maxpool_fwd", ph, pw, float(scale), ptr(features), ptr(rois), ptr(out), ptr(arg), B, C, H, W, N,
             stream())
        ctx.save_for_backward(rois, arg)
        ctx.meta = (B, C, H, W, N, ph, pw, float(scale))
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, dout, _darg):
        rois, arg = ctx.saved_tensors
        B, C, H, W, N, ph, pw, scale = ctx.meta
        dout = f32c(dout)
        g = torch.empty(B, C, H, W, device=dout.device, dtype=torch.float32)
        call("l2s_roi_maxpool_bwd", ph, pw, scale, ptr(dout), ptr(rois), ptr(g), ptr(arg), B, C, H, W, N, stream())
        return g, None, None, None, None


def roi_max_pool(features, rois, pooled_height=7, pooled_width=7, spatial_scale=1.0 / 16, return_argmax=False):
    """RoIPoolFunction (layer_utils/roi_pooling/roi_pool.py:6-50)."""
    out, arg = _RoIMaxPool.apply(features, rois, int(pooled_height), int(pooled_width), float(spatial_scale))
    return (out, arg) if return_argmax else out


# ------------------------------------------------------------------------------------------------
# mask head
# ------------------------------------------------------------------------------------------------
class _MaskHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, up_w, up_b, pred_w, pred_b):
        x, up_w, up_b, pred_w, pred_b = (f32c(t) for t in (x, up_w, up_b, pred_w, pred_b))
        n, Cin = x.shape[0], x.shape[1]
        assert x.shape[2:] == (7, 7), "mask head expects (n,Cin,7,7) res5 features"
        Cmid, ncls = up_w.shape[1], pred_w.shape[0]
        assert up_w.shape == (Cin, Cmid, 2, 2) and pred_w.shape[:2] == (ncls, Cmid)
        score = torch.empty(n, ncls, 14, 14, device=x.device, dtype=torch.float32)
        prob = torch.empty_like(score)
        saved = _ws(_lib.size("l2s_mask_head_saved_bytes", n, Cin, Cmid, ncls), x.device)
        nbytes = _lib.size("l2s_mask_head_workspace_bytes", n, Cin, Cmid, ncls)
        ws = _ws(nbytes, x.device)
        call("l2s_mask_head_fwd", ptr(x), ptr(up_w), ptr(up_b), ptr(pred_w), ptr(pred_b), ptr(score), ptr(prob),
             ptr(saved), n, Cin, Cmid, ncls, ptr(ws), nbytes, stream())
        ctx.save_for_backward(up_w, pred_w, saved, prob)
        ctx.meta = (n, Cin, Cmid, ncls)
        # an unused output (the reference's loss only reads mask_score) must arrive as None in backward, not as a
        # materialised (n,ncls,14,14) tensor of zeros that is then pushed through the sigmoid chain rule
        ctx.set_materialize_grads(False)
        return score, prob

    @staticmethod
    def backward(ctx, dscore, dprob):
        up_w, pred_w, saved, prob = ctx.saved_tensors
        n, Cin, Cmid, ncls = ctx.meta
        if dscore is None:
            dscore = torch.zeros_like(prob)
        dscore = f32c(dscore)
        if dprob is not None:
            dscore = dscore + dprob * prob * (1 - prob)
        dx = torch.empty(n, Cin, 7, 7, device=dscore.device, dtype=torch.float32)
        d_up_w = torch.empty_like(up_w)
        d_up_b = torch.empty(Cmid, device=dscore.device, dtype=torch.float32)
        d_pred_w = torch.empty(ncls, Cmid, device=dscore.device, dtype=torch.float32)
        d_pred_b = torch.empty(ncls, device=dscore.device, dtype=torch.float32)
        nbytes = _lib.size("l2s_mask_head_workspace_bytes", n, Cin, Cmid, ncls)
        ws = _ws(nbytes, dscore.device)
        call("l2s_mask_head_bwd", ptr(dscore), ptr(up_w), ptr(pred_w), ptr(saved), ptr(dx), ptr(d_up_w), ptr(d_up_b),
             ptr(d_pred_w), ptr(d_pred_b), n, Cin, Cmid, ncls, ptr(ws), nbytes, stream())
        return dx, d_up_w, d_up_b, d_pred_w.view_as(pred_w), d_pred_b


def mask_head(x, up_w, up_b, pred_w, pred_b):
    """Network._mask_prediction (network_cycle_response.py:292-307) -> (mask_score, mask_prob)."""
    return _MaskHead.apply(x, up_w, up_b, pred_w, pred_b)


def mask_head_stage_runner(x, up_w, up_b, pred_w, pred_b, stages=2):
    """Runs the full mask-head forward once and returns a closure that re-launches only `stages` of it (1 repacks,
    2 GEMM1 with the bias+ReLU -> bf16-plane epilogue, 4 GEMM2) on the same buffers -- the kernels exactly as the step
    runs them, for timing / profiling."""
    x, up_w, up_b, pred_w, pred_b = (f32c(t.detach()) for t in (x, up_w, up_b, pred_w, pred_b))
    n, Cin = x.shape[0], x.shape[1]
    Cmid, ncls = up_w.shape[1], pred_w.shape[0]
    score = torch.empty(n, ncls, 14, 14, device=x.device, dtype=torch.float32)
    prob = torch.empty_like(score)
    saved = _ws(_lib.size("l2s_mask_head_saved_bytes", n, Cin, Cmid, ncls), x.device)
    nbytes = _lib.size("l2s_mask_head_workspace_bytes", n, Cin, Cmid, ncls)
    ws = _ws(nbytes, x.device)

    def run(st=stages):
        call("l2s_mask_head_fwd_stages", ptr(x), ptr(up_w), ptr(up_b), ptr(pred_w), ptr(pred_b), ptr(score), ptr(prob),
             ptr(saved), n, Cin, Cmid, ncls, ptr(ws), nbytes, st, stream())

    run(7)
    run.keep = (x, up_w, up_b, pred_w, pred_b, score, prob, saved, ws)
    return run


class _MaskBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, labels, target):
        score, target = f32c(score), f32c(target)
        labels = labels.to(torch.int64).contiguous()
        n, ncls = score.shape[:2]
        hw = score.shape[2] * score.shape[3]
        loss = torch.empty(1, device=score.device, dtype=torch.float32)
        call("l2s_mask_bce_fwd", ptr(score), ptr(labels), ptr(target), ptr(loss), n, ncls, hw, stream())
        ctx.save_for_backward(score, labels, target)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        score, labels, target = ctx.saved_tensors
        n, ncls = score.shape[:2]
        hw = score.shape[2] * score.shape[3]
        dscore = torch.empty_like(score)
        gs = f32c(g).reshape(1)
        call("l2s_mask_bce_bwd", ptr(score), ptr(labels), ptr(target), ptr(gs), ptr(dscore), n, ncls, hw, stream())
        return dscore, None, None


def mask_targets(gt_masks, rois, gt_assignment, size=14):
    """Mask targets of the proposal target layer (proposal_target_layer.py:193-201) on the device.

    gt_masks (G,imH,imW) uint8 {0,1} ; rois (n,5) [batch,x1,y1,x2,y2] image pixels ; gt_assignment (n) ->
    (n,size,size) float {0,1}: masks[assign[i], int(y1):int(y2)+1, int(x1):int(x2)+1] resized 'nearest'."""
    assert gt_masks.is_cuda and gt_masks.dtype == torch.uint8 and gt_masks.dim() == 3
    m = gt_masks.contiguous()
    r = f32c(rois)
    a = gt_assignment.to(device=m.device, dtype=torch.int32).contiguous()
    n = r.shape[0]
    out = torch.empty(n, size, size, device=m.device, dtype=torch.float32)
    call("l2s_mask_crop_resize", ptr(m), ptr(r), r.shape[1], ptr(a), ptr(out), m.shape[0], m.shape[1], m.shape[2], n,
         size, size, stream())
    return out


def resize_masks_nearest(gt_masks, H, W):
    """imresize(mask, (H,W), interp='nearest') of network_cycle_response.py:418 for a stack of uint8 {0,1} masks
    (G,imH,imW) -> (G,H,W) float: the response target, built on the device."""
    assert gt_masks.is_cuda and gt_masks.dtype == torch.uint8 and gt_masks.dim() == 3
    m = gt_masks.contiguous()
    out = torch.empty(m.shape[0], H, W, device=m.device, dtype=torch.float32)
    call("l2s_mask_crop_resize", ptr(m), None, 0, None, ptr(out), m.shape[0], m.shape[1], m.shape[2], m.shape[0], H, W,
         stream())
    return out


def colsum(x):
    """Sum over every dimension but the last (bias gradients) of an fp32 CUDA tensor, fixed summation order.
    Accepts a row-strided 2-D view (stride(1) == 1) without copying; anything else is made contiguous first."""
    if x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32:
        R, C, ld = x.shape[0], x.shape[1], x.stride(0)
    else:
        x = f32c(x)
        C = x.shape[-1]
        R, ld = x.numel() // C, C
    out = torch.empty(C, device=x.device, dtype=torch.float32)
    nbytes = _lib.size("l2s_colsum_workspace_bytes", R, C)
    ws = _ws(nbytes, x.device)
    call("l2s_colsum", ptr(x), ld, ptr(out), R, C, ptr(ws), nbytes, stream())
    return out


def nms_sorted(boxes_sorted, thresh, max_out=0):
    """gpu_nms (nms_cuda.c:17-67) on score-sorted boxes (N,5), scan included, on the device ->
    (keep (N,) int64: positions of the survivors in score order, valid up to num ; num (1,) int64)."""
    b = f32c(boxes_sorted)
    assert b.dim() == 2 and b.shape[1] == 5
    n = b.shape[0]
    keep = torch.zeros(max(n, 1), device=b.device, dtype=torch.int64)
    num = torch.zeros(1, device=b.device, dtype=torch.int64)
    nbytes = _lib.size("l2s_nms_workspace_bytes", n)
    ws = _ws(nbytes, b.device)
    call("l2s_nms", ptr(b), n, float(thresh), int(max_out), ptr(keep), ptr(num), ptr(ws), nbytes, stream())
    return keep[:n], num


def _bce_du_supported(Cmid):
    return Cmid % 8 == 0 and Cmid // 8 <= 256 and 256 % (Cmid // 8) == 0


class _MaskHeadLoss(torch.autograd.Function):
    """Mask head + mask loss as ONE autograd node.  When the loss is the only consumer of the scores (the reference's
    training step) the backward never materialises dscore: l2s_mask_head_bce_bwd builds the dU planes and the
    prediction-layer gradients straight from (score, labels, target).  Any other upstream gradient on score / prob
    falls back to the general path (l2s_mask_bce_bwd + l2s_mask_head_bwd)."""

    @staticmethod
    def forward(ctx, x, up_w, up_b, pred_w, pred_b, labels, target):
        x, up_w, up_b, pred_w, pred_b, target = (f32c(t) for t in (x, up_w, up_b, pred_w, pred_b, target))
        labels = labels.to(torch.int64).contiguous()
        n, Cin = x.shape[0], x.shape[1]
        assert x.shape[2:] == (7, 7), "mask head expects (n,Cin,7,7) res5 features"
        Cmid, ncls = up_w.shape[1], pred_w.shape[0]
        assert up_w.shape == (Cin, Cmid, 2, 2) and pred_w.shape[:2] == (ncls, Cmid)
        score = torch.empty(n, ncls, 14, 14, device=x.device, dtype=torch.float32)
        prob = torch.empty_like(score)
        saved = _ws(_lib.size("l2s_mask_head_saved_bytes", n, Cin, Cmid, ncls), x.device)
        nbytes = _lib.size("l2s_mask_head_workspace_bytes", n, Cin, Cmid, ncls)
        ws = _ws(nbytes, x.device)
        call("l2s_mask_head_fwd", ptr(x), ptr(up_w), ptr(up_b), ptr(pred_w), ptr(pred_b), ptr(score), ptr(prob),
             ptr(saved), n, Cin, Cmid, ncls, ptr(ws), nbytes, stream())
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        call("l2s_mask_bce_fwd", ptr(score), ptr(labels), ptr(target), ptr(loss), n, ncls, 196, stream())
        ctx.save_for_backward(up_w, pred_w, saved, prob, score, labels, target)
        ctx.meta = (n, Cin, Cmid, ncls)
        ctx.set_materialize_grads(False)
        return score, prob, loss[0]

    @staticmethod
    def backward(ctx, dscore, dprob, dloss):
        up_w, pred_w, saved, prob, score, labels, target = ctx.saved_tensors
        n, Cin, Cmid, ncls = ctx.meta
        dev = score.device
        dx = torch.empty(n, Cin, 7, 7, device=dev, dtype=torch.float32)
        d_up_w = torch.empty_like(up_w)
        d_up_b = torch.empty(Cmid, device=dev, dtype=torch.float32)
        d_pred_w = torch.empty(ncls, Cmid, device=dev, dtype=torch.float32)
        d_pred_b = torch.empty(ncls, device=dev, dtype=torch.float32)
        nbytes = _lib.size("l2s_mask_head_workspace_bytes", n, Cin, Cmid, ncls)
        ws = _ws(nbytes, dev)
        if dscore is None and dprob is None and _bce_du_supported(Cmid):
            gs = f32c(dloss).reshape(1) if dloss is not None else torch.zeros(1, device=dev)
            call("l2s_mask_head_bce_bwd", ptr(score), ptr(labels), ptr(target), ptr(gs), ptr(up_w), ptr(pred_w),
                 ptr(saved), ptr(dx), ptr(d_up_w), ptr(d_up_b), ptr(d_pred_w), ptr(d_pred_b), n, Cin, Cmid, ncls,
                 ptr(ws), nbytes, stream())
        else:
            ds = torch.empty_like(score)
            gs = f32c(dloss).reshape(1) if dloss is not None else torch.zeros(1, device=dev)
            call("l2s_mask_bce_bwd", ptr(score), ptr(labels), ptr(target), ptr(gs), ptr(ds), n, ncls, 196, stream())
            if dscore is not None:
                ds = ds + f32c(dscore)
            if dprob is not None:
                ds = ds + dprob * prob * (1 - prob)
            call("l2s_mask_head_bwd", ptr(ds), ptr(up_w), ptr(pred_w), ptr(saved), ptr(dx), ptr(d_up_w), ptr(d_up_b),
                 ptr(d_pred_w), ptr(d_pred_b), n, Cin, Cmid, ncls, ptr(ws), nbytes, stream())
        return dx, d_up_w, d_up_b, d_pred_w.view_as(pred_w), d_pred_b, None, None


def mask_head_with_loss(x, up_w, up_b, pred_w, pred_b, labels, mask_targets):
    """_mask_prediction (network_cycle_response.py:292-307) and the mask loss (:404-413) in one node ->
    (mask_score, mask_prob, loss).  Same values as mask_head(...) followed by mask_bce_loss(...)."""
    return _MaskHeadLoss.apply(x, up_w, up_b, pred_w, pred_b, labels, mask_targets)


def mask_bce_loss(mask_score, labels, mask_targets):
    """Mask loss of network_cycle_response.py:404-413."""
    return _MaskBCE.apply(mask_score, labels, mask_targets)


# ------------------------------------------------------------------------------------------------
# box head glue (row a9)
# ------------------------------------------------------------------------------------------------
class _SpatialMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = f32c(x)
        N, C = x.shape[0], x.shape[1]
        P = x.numel() // max(N * C, 1)
        out = torch.empty(N, C, device=x.device, dtype=torch.float32)
        call("l2s_spatial_mean_fwd", ptr(x), ptr(out), N * C, P, stream())
        ctx.shape = tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = f32c(dout)
        dx = torch.empty(ctx.shape, device=dout.device, dtype=torch.float32)
        N, C = ctx.shape[0], ctx.shape[1]
        call("l2s_spatial_mean_bwd", ptr(dout), ptr(dx), N * C, dx.numel() // max(N * C, 1), stream())
        return dx


def spatial_mean(x):
    """spatial_fc7.mean(3).mean(2) (network_cycle_response.py:278): (N,C,h,w) -> (N,C)."""
    return _SpatialMean.apply(x)


def softmax_argmax(score):
    """(F.softmax(score, 1), torch.max(score, 1)[1]) of network_cycle_response.py:280-281 in one kernel (no gradient:
    the reference's losses read cls_score, not cls_prob).  `score` may be a column slice of a wider matrix."""
    assert score.dim() == 2 and score.stride(1) == 1 and score.dtype == torch.float32
    R, ncls = score.shape
    prob = torch.empty(R, ncls, device=score.device, dtype=torch.float32)
    pred = torch.empty(R, device=score.device, dtype=torch.int64)
    call("l2s_softmax_argmax", ptr(score), score.stride(0) if R > 1 else ncls, ptr(prob), ptr(pred), R, ncls, stream())
    return prob, pred


def region_classification(spatial_fc7, cls_w, cls_b, bbox_w, bbox_b):
    """Network._region_classification (network_cycle_response.py:277-290): 7x7 mean -> cls_score_net and bbox_pred_net
    as ONE stacked GEMM (tcgen05 bf16x3 for >= 512 ROIs, the skinny exact-fp32 GEMM below that) -> softmax / argmax.
    Returns (cls_score, cls_pred, cls_prob, bbox_pred)."""
    fc7 = spatial_mean(spatial_fc7)
    ncls, nbox = cls_w.shape[0], bbox_w.shape[0]
    pad = (-(ncls + nbox)) % 8                       # operand rows of the tensor-core GEMM must be 16-byte multiples
    W = torch.cat([cls_w, bbox_w] + ([cls_w.new_zeros(pad, cls_w.shape[1])] if pad else []), 0)
    b = torch.cat([cls_b, bbox_b] + ([cls_b.new_zeros(pad)] if pad else []), 0)
    y = linear(fc7, W, b) if fc7.shape[0] >= 512 and fc7.shape[1] % 8 == 0 else linear_small_fn(fc7, W, b)
    cls_score, bbox_pred = y[:, :ncls], y[:, ncls:ncls + nbox]
    cls_prob, cls_pred = softmax_argmax(cls_score.detach())
    return cls_score, cls_pred, cls_prob, bbox_pred


# ------------------------------------------------------------------------------------------------
# (3) attention step and the att2in2 epilogues
# ------------------------------------------------------------------------------------------------
class _AttStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, att_h, att_feats, p_att, alpha_w, alpha_b):
        att_h, att_feats, p_att = f32c(att_h), f32c(att_feats), f32c(p_att)
        ctx.w_shape, ctx.b_shape = alpha_w.shape, alpha_b.shape
        alpha_w, alpha_b = f32c(alpha_w).view(-1), f32c(alpha_b).view(-1)
        B, A, D = att_feats.shape
        Dh = p_att.shape[2]
        weight = torch.empty(B, A, device=att_h.device, dtype=torch.float32)
        res = torch.empty(B, D, device=att_h.device, dtype=torch.float32)
        call("l2s_att_step_fwd", ptr(att_h), ptr(att_feats), ptr(p_att), ptr(alpha_w), ptr(alpha_b), ptr(weight),
             ptr(res), B, A, D, Dh, stream())
        ctx.save_for_backward(att_h, att_feats, p_att, alpha_w, weight)
        return res, weight

    @staticmethod
    def backward(ctx, dres, _dweight):
        att_h, att_feats, p_att, alpha_w, weight = ctx.saved_tensors
        B, A, D = att_feats.shape
        Dh = p_att.shape[2]
        dres = f32c(dres)
        datt_h = torch.empty_like(att_h)
        de = torch.empty_like(weight)
        dp_att = torch.zeros_like(p_att)
        datt = torch.zeros_like(att_feats)
        dalpha = torch.zeros_like(alpha_w)
        call("l2s_att_step_bwd", ptr(dres), ptr(att_h), ptr(att_feats), ptr(p_att), ptr(alpha_w), ptr(weight),
             ptr(datt_h), ptr(de), ptr(dp_att), ptr(datt), ptr(dalpha), B, A, D, Dh, stream())
        return datt_h, datt, dp_att, dalpha.view(ctx.w_shape), de.sum().reshape(ctx.b_shape)


def attention_step(att_h, att_feats, p_att, alpha_w, alpha_b):
    """Attention.forward after h2att (AttModel.py:411-421) -> (att_res (B,D), weight (B,A))."""
    return _AttStep.apply(att_h, att_feats, p_att, alpha_w, alpha_b)


class _Gates(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sums, a2c_out, c_prev):
        sums, a2c_out, c_prev = f32c(sums), f32c(a2c_out), f32c(c_prev)
        B, D = c_prev.shape
        h = torch.empty_like(c_prev)
        c = torch.empty_like(c_prev)
        call("l2s_att2in2_gates_fwd", ptr(sums), ptr(a2c_out), ptr(c_prev), ptr(h), ptr(c), B, D, stream())
        ctx.save_for_backward(sums, a2c_out, c_prev, c)
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        sums, a2c_out, c_prev, c = ctx.saved_tensors
        B, D = c_prev.shape
        dh = f32c(dh) if dh is not None else torch.zeros_like(c)
        dc = f32c(dc) if dc is not None else None
        dsums = torch.empty_like(sums)
        da2c = torch.empty_like(a2c_out)
        dcp = torch.empty_like(c_prev)
        call("l2s_att2in2_gates_bwd", ptr(sums), ptr(a2c_out), ptr(c_prev), ptr(c), ptr(dh), ptr(dc), ptr(dsums),
             ptr(da2c), ptr(dcp), B, D, stream())
        return dsums, da2c, dcp


def att2in2_gates(sums, a2c_out, c_prev):
    """Att2in2Core gate epilogue (AttModel.py:450-462) -> (next_h, next_c)."""
    return _Gates.apply(sums, a2c_out, c_prev)


class _Att2in2Decode(torch.autograd.Function):
    """The whole T-step att2in2 recurrence (AttModel.py:75-99 around Att2in2Core :446-466) as one call per
    direction: l2s_att2in2_decode_{fwd,bwd}.  Weight gradients are GEMMs over the stacked T*B rows."""

    @staticmethod
    def forward(ctx, i2h_all, att, p_att, w_h2att, b_h2att, w_h2h, w_a2c, b_a2c, alpha_w, alpha_b):
        i2h_all, att, p_att = f32c(i2h_all), f32c(att), f32c(p_att)
        T, B, _ = i2h_all.shape
        A, D = att.shape[1], att.shape[2]
        Dh = p_att.shape[2]
        LC = Dh + 5 * D
        dev = att.device
        ctx.shapes = (T, B, A, D, Dh, alpha_w.shape, alpha_b.shape)
        cat_all = torch.empty(T, B, LC, device=dev, dtype=torch.float32)
        cat_all[:, :, :Dh] = f32c(b_h2att)
        cat_all[:, :, Dh:] = i2h_all
        w_cat = torch.cat([f32c(w_h2att), f32c(w_h2h)], 0).contiguous()
        w_a2c, b_a2c = f32c(w_a2c), f32c(b_a2c)
        aw, ab = f32c(alpha_w).reshape(-1), f32c(alpha_b).reshape(-1)
        h_all = torch.empty(T, B, D, device=dev, dtype=torch.float32)
        c_all = torch.empty_like(h_all)
        a2c_all = torch.empty(T, B, 2 * D, device=dev, dtype=torch.float32)
        pi_all = torch.empty(T, B, A, device=dev, dtype=torch.float32)
        res_all = torch.empty_like(h_all)
        nbytes = _lib.size("l2s_att2in2_decode_workspace_bytes", T, B, A, D, Dh)
        ws = _ws(nbytes, dev)
        call("l2s_att2in2_decode_fwd", ptr(cat_all), ptr(att), ptr(p_att), ptr(w_cat), ptr(w_a2c), ptr(b_a2c), ptr(aw),
             ptr(ab), ptr(h_all), ptr(c_all), ptr(a2c_all), ptr(pi_all), ptr(res_all), T, B, A, D, Dh, ptr(ws), nbytes,
             stream())
        ctx.save_for_backward(cat_all, att, p_att, w_cat, w_a2c, aw, h_all, c_all, a2c_all, pi_all, res_all)
        return h_all

    @staticmethod
    def backward(ctx, dh_all):
        cat_all, att, p_att, w_cat, w_a2c, aw, h_all, c_all, a2c_all, pi_all, res_all = ctx.saved_tensors
        T, B, A, D, Dh, aw_shape, ab_shape = ctx.shapes
        LC = Dh + 5 * D
        dev = att.device
        dh_all = f32c(dh_all)
        w_cat_t = w_cat.t().contiguous()
        w_a2c_t = w_a2c.t().contiguous()
        dcat_all = torch.empty(T, B, LC, device=dev, dtype=torch.float32)
        da2c_all = torch.empty(T, B, 2 * D, device=dev, dtype=torch.float32)
        dres_all = torch.empty(T, B, D, device=dev, dtype=torch.float32)
        de_all = torch.empty(T, B, A, device=dev, dtype=torch.float32)
        dp_att = torch.empty_like(p_att)
        datt = torch.empty_like(att)
        dalpha = torch.empty(Dh, device=dev, dtype=torch.float32)
        nbytes = _lib.size("l2s_att2in2_decode_workspace_bytes", T, B, A, D, Dh)
        ws = _ws(nbytes, dev)
        call("l2s_att2in2_decode_bwd", ptr(dh_all), ptr(cat_all), ptr(att), ptr(p_att), ptr(w_cat_t), ptr(w_a2c_t),
             ptr(aw), ptr(c_all), ptr(a2c_all), ptr(pi_all), ptr(dcat_all), ptr(da2c_all), ptr(dres_all), ptr(de_all),
             ptr(dp_att), ptr(datt), ptr(dalpha), T, B, A, D, Dh, ptr(ws), nbytes, stream())
        # weight gradients: one GEMM each over the stacked rows (h_{-1} = 0 contributes nothing)
        dw_cat = wgrad(dcat_all[1:].reshape(-1, LC), h_all[:-1].reshape(-1, D))
        dw_a2c = wgrad(da2c_all.reshape(-1, 2 * D), res_all.reshape(-1, D))
        return (dcat_all[:, :, Dh:], datt, dp_att, dw_cat[:Dh], colsum(dcat_all.view(T * B, LC)[:, :Dh]), dw_cat[Dh:], dw_a2c,
                colsum(da2c_all), dalpha.view(aw_shape), colsum(de_all.view(T * B, -1)).sum().reshape(ab_shape))


def att2in2_decode(i2h_all, att_feats, p_att, w_h2att, b_h2att, w_h2h, w_a2c, b_a2c, alpha_w, alpha_b):
    """h_t for t = 0..T-1 of the att2in2 recurrence from zero state.

    i2h_all (T,B,5D) = i2h(x_t) + b_i2h + b_h2h ; att_feats (B,A,D) ; p_att (B,A,Dh), Dh == D -> h_all (T,B,D)."""
    return _Att2in2Decode.apply(i2h_all, att_feats, p_att, w_h2att, b_h2att, w_h2h, w_a2c, b_a2c, alpha_w, alpha_b)


class _LinearSmallFn(torch.autograd.Function):
    """nn.Linear on a small batch (rows <= a few hundred) in exact fp32 through l2s_linear_small, fwd and dX (the
    strided FFMA GEMM l2s_gemm_f32 where a reduction length is not a multiple of 4); the weight gradient is one
    rank-`rows` update (cuBLAS)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = f32c(x), f32c(w)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        if x.shape[1] % 4 == 0:
            return linear_small(x, w, b)
        y = gemm_f32(x, w)
        return y + f32c(b) if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = f32c(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            M, N = dy.shape
            K = w.shape[1]
            if N % 4 == 0:
                dx = linear_small(dy, w.t().contiguous())
            else:       # dx[m,k] = sum_n dy[m,n] w[n,k]: w read in place with strides (1, K)
                dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
                call("l2s_gemm_f32", ptr(dy), ptr(w), ptr(dx), M, K, N, N, 1, 1, K, K, 0, stream())
        if ctx.needs_input_grad[1]:
            dw = wgrad_f32(dy, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dx, dw, db


def linear_small_fn(x, weight, bias=None):
    """Differentiable x @ weight.T + bias for small row counts (exact fp32, no cuBLAS SIMT launches)."""
    return _LinearSmallFn.apply(x, weight, bias)


def linear_small(x, weight, bias=None, out=None, accumulate=False):
    """out (+)= x @ weight.T + bias, exact fp32, for a small number of rows (no gradient)."""
    x, weight = f32c(x), f32c(weight)
    M, K = x.shape
    N = weight.shape[0]
    D = out if out is not None else torch.empty(M, N, device=x.device, dtype=torch.float32)
    nbytes = _lib.size("l2s_linear_small_workspace_bytes", M, N, K)
    ws = _ws(nbytes, x.device)
    call("l2s_linear_small", ptr(x), ptr(weight), ptr(f32c(bias) if bias is not None else None), ptr(D), M, N, K, K, K,
         D.stride(0), int(accumulate), ptr(ws), nbytes, stream())
    return D


class _BiLSTM(torch.autograd.Function):
    """Masked variable-length bidirectional LSTM recurrence (l2s_bilstm_{fwd,bwd}); the input projection and the
    weight gradients are GEMMs over all (b,t) rows outside the time loop."""

    @staticmethod
    def forward(ctx, xg, w_hh_f, w_hh_b, lens):
        B, L, H8 = xg.shape
        H = H8 // 8
        dev = xg.device
        G = f32c(xg).clone()
        w_hh = torch.stack([f32c(w_hh_f), f32c(w_hh_b)], 0).contiguous()
        lens = lens.to(device=dev, dtype=torch.int32).contiguous()
        c_all = torch.empty(L, 2, B, H, device=dev, dtype=torch.float32)
        h_all = torch.empty_like(c_all)
        out = torch.empty(B, L, 2 * H, device=dev, dtype=torch.float32)
        hidden = torch.empty(B, 2 * H, device=dev, dtype=torch.float32)
        nbytes = _lib.size("l2s_bilstm_workspace_bytes", L, B, H)
        ws = _ws(nbytes, dev)
        call("l2s_bilstm_fwd", ptr(G), ptr(w_hh), ptr(lens), ptr(c_all), ptr(h_all), ptr(out), ptr(hidden), L, B, H,
             ptr(ws), nbytes, stream())
        ctx.save_for_backward(G, w_hh, lens, c_all, h_all)
        ctx.set_materialize_grads(False)       # `out` is usually unused: its gradient arrives as None (NULL dout)
        return out, hidden

    @staticmethod
    def backward(ctx, dout, dhidden):
        G, w_hh, lens, c_all, h_all = ctx.saved_tensors
        B, L, H8 = G.shape
        H = H8 // 8
        dev = G.device
        dout = f32c(dout) if dout is not None else None
        dhidden = f32c(dhidden) if dhidden is not None else None
        if dout is None and dhidden is None:
            return torch.zeros_like(G), torch.zeros_like(w_hh[0]), torch.zeros_like(w_hh[1]), None
        w_hh_t = w_hh.transpose(1, 2).contiguous()
        dG = torch.empty_like(G)
        nbytes = _lib.size("l2s_bilstm_workspace_bytes", L, B, H)
        ws = _ws(nbytes, dev)
        call("l2s_bilstm_bwd", ptr(dout), ptr(dhidden), ptr(G), ptr(w_hh_t), ptr(lens), ptr(c_all), ptr(dG), L, B, H,
             ptr(ws), nbytes, stream())
        # recurrent weight gradients: one GEMM per direction over all (t,b) rows against the previous state
        dG5 = dG.view(B, L, 2, 4 * H)
        zero = h_all.new_zeros(1, B, H)
        hp_f = torch.cat([zero, h_all[:-1, 0]], 0)           # state before time t, forward direction
        hp_b = torch.cat([h_all[1:, 1], zero], 0)            # ... backward direction (previous = time t+1)
        dw_f = wgrad(dG5[:, :, 0].transpose(0, 1).reshape(L * B, 4 * H), hp_f.reshape(L * B, H))
        dw_b = wgrad(dG5[:, :, 1].transpose(0, 1).reshape(L * B, 4 * H), hp_b.reshape(L * B, H))
        return dG, dw_f, dw_b, None


def bilstm(xg, w_hh_f, w_hh_b, lens):
    """xg (B,L,8H) = x W_ih^T + b for [forward | backward] direction (gate order i,f,g,o) ; lens (B) ->
    (out (B,L,2H) zero padded, hidden (B,2H) = [h_fwd(last) | h_bwd(first)])."""
    return _BiLSTM.apply(xg, w_hh_f, w_hh_b, lens)


class _LogSoftmaxNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, mask, want_logp):
        logits = f32c(logits)
        R, V = logits.shape
        target = target.to(torch.int64).contiguous()
        mask = f32c(mask)
        logp = torch.empty_like(logits) if want_logp else None
        nll = torch.empty(R, device=logits.device, dtype=torch.float32)
        call("l2s_logsoftmax_nll_fwd", ptr(logits), ptr(target), ptr(mask), ptr(logp), ptr(nll), R, V, stream())
        ctx.save_for_backward(logits, target, mask)
        if logp is None:
            logp = logits.new_empty(0)
        ctx.mark_non_differentiable(logp)
        return nll.sum(), logp

    @staticmethod
    def backward(ctx, g, _glogp):
        logits, target, mask = ctx.saved_tensors
        R, V = logits.shape
        d = torch.empty_like(logits)
        gs = f32c(g).reshape(1)
        call("l2s_logsoftmax_nll_bwd", ptr(logits), ptr(target), ptr(mask), ptr(gs), ptr(d), R, V, stream())
        return d, None, None, None


def logsoftmax_nll(logits, target, mask, want_logp=False):
    """sum_r -log_softmax(logits)[r,target[r]] * mask[r]  (AttModel.py:98 + misc/utils.py:43-53 numerator)."""
    return _LogSoftmaxNLL.apply(logits, target, mask, want_logp)


def log_softmax(logits):
    """log_softmax over the last dim of (R,V) logits through the fused kernel (no gradient)."""
    logits = f32c(logits)
    R, V = logits.shape
    logp = torch.empty_like(logits)
    call("l2s_logsoftmax_nll_fwd", ptr(logits), None, None, ptr(logp), None, R, V, stream())
    return logp


# ------------------------------------------------------------------------------------------------
# caption feature prep
# ------------------------------------------------------------------------------------------------
class _CaptionFeats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, before, after, S):
        before, after = f32c(before), f32c(after)
        B, C, H, W = before.shape
        fc = torch.empty(B, 2 * C, device=before.device, dtype=torch.float32)
        att = torch.empty(B, S, S, 2 * C, device=before.device, dtype=torch.float32)
        call("l2s_caption_feats_fwd", ptr(before), ptr(fc), ptr(att), B, C, H, W, S, 2 * C, 0, stream())
        call("l2s_caption_feats_fwd", ptr(after), ptr(fc), ptr(att), B, C, H, W, S, 2 * C, C, stream())
        ctx.meta = (B, C, H, W, S)
        return fc, att

    @staticmethod
    def backward(ctx, dfc, datt):
        B, C, H, W, S = ctx.meta
        dev = (dfc if dfc is not None else datt).device
        dfc = f32c(dfc) if dfc is not None else None
        datt = f32c(datt) if datt is not None else None
        gb = torch.empty(B, C, H, W, device=dev, dtype=torch.float32)
        ga = torch.empty(B, C, H, W, device=dev, dtype=torch.float32)
        call("l2s_caption_feats_bwd", ptr(dfc), ptr(datt), ptr(gb), B, C, H, W, S, 2 * C, 0, stream())
        call("l2s_caption_feats_bwd", ptr(dfc), ptr(datt), ptr(ga), B, C, H, W, S, 2 * C, C, stream())
        return gb, ga, None


def caption_features(feats_before, feats_after, att_size=14):
    """network_cycle_response.py:428-438 -> fc (B,2C), att (B,S,S,2C)."""
    return _CaptionFeats.apply(feats_before, feats_after, att_size)


# ------------------------------------------------------------------------------------------------
# precision of the tensor-core GEMMs (mask head, caption projections)
# ------------------------------------------------------------------------------------------------
_PRECISIONS = {"fp32": 0, "bf16": 1}


def set_precision(mode):
    """'fp32' (default): bf16x3 split products, ~1e-5 from an fp32 GEMM (the 1e-4 parity contract).
    'bf16': one tensor pass on the bf16-rounded operands -- north_star's bf16 variants (1e-2).  Process wide;
    a CUDA graph keeps the mode it was captured with."""
    if mode not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
    call("l2s_set_precision", _PRECISIONS[mode])


def get_precision():
    return {v: k for k, v in _PRECISIONS.items()}[int(_lib.load().l2s_get_precision())]


class precision:
    """with precision('bf16'): ...  -- scoped set_precision()."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)
        return self

    def __exit__(self, *exc):
        set_precision(self.prev)
        return False


# ------------------------------------------------------------------------------------------------
# raw GEMM access (tests / benchmarks)
# ------------------------------------------------------------------------------------------------
def split_bf16(x, ld_dst=None):
    x = f32c(x)
    rows, cols = x.shape
    ld = ld_dst or cols
    hi = torch.empty(rows, ld, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty(rows, ld, device=x.device, dtype=torch.bfloat16)
    call("l2s_split_bf16", ptr(x), ptr(hi), ptr(lo), rows, cols, cols, ld, stream())
    return hi, lo


def gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn=False, b_mn=False, epilogue=0, bias=None, bias_div=1,
                split_k=1, out=None):
    D = out if out is not None else torch.empty(M, N, device=a_hi.device, dtype=torch.float32)
    call("l2s_gemm_bf16x3", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), ptr(D), ptr(bias), bias_div, M, N, K,
         int(a_mn), int(b_mn), epilogue, split_k, stream())
    return D


class _LinearTC(torch.autograd.Function):
    """y = x @ w.T + b on the tcgen05 bf16x3 GEMM; both gradients read the saved bf16 planes in place
    through MN-major descriptors (no transposes)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = f32c(x), f32c(w)
        M, K = x.shape
        N = w.shape[0]
        xh, xl = split_bf16(x)
        wh, wl = split_bf16(w)
        bias = f32c(b) if b is not None else None
        y = gemm_bf16x3(xh, xl, wh, wl, M, N, K, epilogue=4 if bias is not None else 0, bias=bias)
        ctx.save_for_backward(xh, xl, wh, wl)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xh, xl, wh, wl = ctx.saved_tensors
        M, K = xh.shape
        N = wh.shape[0]
        dy = f32c(dy)
        dyh, dyl = split_bf16(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dY [M,N] . W [N,K]; with a few hundred rows (i2h, logit: T*B = 528) there are only 8-10 output tiles for
            # 148 SMs and each walks the whole N: let the library split the contraction (as for dW)
            few_tiles = ((M + 127) // 128) * ((K + 255) // 256) * 2 <= 148 and N >= 1024
            dx = gemm_bf16x3(dyh, dyl, wh, wl, M, K, N, a_mn=False, b_mn=True, split_k=0 if few_tiles else 1)
        if ctx.needs_input_grad[1]:
            dw = gemm_bf16x3(dyh, dyl, xh, xl, N, K, M, a_mn=True, b_mn=True, split_k=0)  # dY^T [N,M] . X [M,K]; split chosen by the library
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dx, dw, db


def linear(x, weight, bias=None):
    """nn.Linear forward/backward for large row counts on the tensor pipe at fp32 accuracy (bf16x3)."""
    return _LinearTC.apply(x, weight, bias)


def dense(x, weight, bias=None):
    """nn.Linear on the library's own kernels, chosen by shape: the tcgen05 bf16x3 GEMM (fp32 accuracy, ~1e-5) when
    there are enough rows to fill 128-row MMA tiles and the operands meet TMA's 16-byte row alignment, the skinny
    exact-fp32 GEMM otherwise.  There is no cuBLAS / CPU path behind it: CPU tensors raise."""
    if not x.is_cuda:
        raise _lib.L2SError("lang2seg_b200.dense: CUDA tensors only (there is no CPU path)")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    M, K = x2.shape
    N = weight.shape[0]
    # few rows but a large weight (filter generator: 48 expressions x (7C + 8) x 1024): the skinny FFMA kernels then
    # spend 46 (fwd) + 79 (dX incl. the transposed copy of W) + 55 (dW) us per step on work the tensor pipe does in ~10 us
    # each -- the mostly empty 128-row tiles do not matter, the weight traffic does
    if (M >= 256 or (M >= 8 and N * K >= (1 << 21) and not _DENSE_SMALL_FFMA)) and K % 8 == 0 and N % 8 == 0:
        y = linear(x2, weight, bias)
    else:
        y = linear_small_fn(x2, weight, bias)
    return y.view(*lead, N)


def wgrad_f32(dy, x):
    """dW = dy^T @ x for a small number of rows (dy (R,N), x (R,K) -> (N,K)), exact fp32 on the library's FFMA GEMM:
    both operands are read in place with unit row stride (no transposes, no cuBLAS SIMT launch)."""
    dy, x = f32c(dy), f32c(x)
    R, N = dy.shape
    K = x.shape[1]
    D = torch.empty(N, K, device=dy.device, dtype=torch.float32)
    call("l2s_gemm_f32", ptr(dy), ptr(x), ptr(D), N, K, R, 1, N, 1, K, K, 0, stream())
    return D


class _Embedding(torch.autograd.Function):
    """nn.Embedding lookup whose weight gradient is the library's fixed-order row sum (l2s_embedding_bwd) instead of
    torch's sort + segmented reduce (24-30 us per call for the ~500 tokens of a step)."""

    @staticmethod
    def forward(ctx, idx, weight):
        ctx.save_for_backward(idx)
        ctx.vd = weight.shape
        return weight.index_select(0, idx.reshape(-1)).view(*idx.shape, weight.shape[1])

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        V, D = ctx.vd
        dy = f32c(dy).view(-1, D)
        dW = torch.empty(V, D, device=dy.device, dtype=torch.float32)
        call("l2s_embedding_bwd", ptr(idx.reshape(-1).contiguous()), ptr(dy), ptr(dW), dy.shape[0], V, D, stream())
        return None, dW


def embedding(idx, weight):
    """weight[idx] (idx int64 of any shape, weight (V,D) fp32 with D % 4 == 0)."""
    return _Embedding.apply(idx, weight)


class Embedding(torch.nn.Embedding):
    """nn.Embedding with the same parameter name / shape (checkpoints load unchanged); CUDA training lookups go through
    `embedding` above, everything else (CPU tensors, padding_idx / max_norm / sparse options) through torch."""

    def forward(self, input):
        if (input.is_cuda and self.weight.is_cuda and input.dtype == torch.int64 and self.padding_idx is None
                and self.max_norm is None and not self.sparse and not self.scale_grad_by_freq
                and self.weight.dtype == torch.float32 and self.weight.shape[1] % 4 == 0
                and torch.is_grad_enabled() and self.weight.requires_grad):
            return embedding(input, self.weight)
        return super().forward(input)


def wgrad(dy, x):
    """dW = dy^T @ x (dy (R,N), x (R,K) -> (N,K)) for the recurrences' weights.  With a few hundred stacked rows (T*B of
    the decode loop, L*B of the bi-LSTM) the FFMA GEMM above runs for ~55 us per call at a quarter of the fp32 peak; the
    tcgen05 bf16x3 GEMM reads the same operands in place through MN-major descriptors (no transposes) and needs no
    split-K here (16-48 output tiles, 8-9 k-blocks each), so the result stays bit-reproducible.  L2S_WGRAD_FFMA=1 keeps
    the FFMA kernel (A/B)."""
    R, N = dy.shape
    K = x.shape[1]
    if R >= 256 and N % 8 == 0 and K % 8 == 0 and not _WGRAD_FFMA:
        dyh, dyl = split_bf16(dy)
        xh, xl = split_bf16(x)
        return gemm_bf16x3(dyh, dyl, xh, xl, N, K, R, a_mn=True, b_mn=True, split_k=1)
    return wgrad_f32(dy, x)


_WGRAD_FFMA = os.environ.get("L2S_WGRAD_FFMA", "0") == "1"
_DENSE_SMALL_FFMA = os.environ.get("L2S_DENSE_SMALL_FFMA", "0") == "1"      # A/B: skinny FFMA kernels for few-row linears


def gemm_f32(A, B, accumulate=False, out=None):
    """D = A @ B.T with A (M,K), B (N,K), exact fp32 FFMA."""
    A, B = f32c(A), f32c(B)
    M, K = A.shape
    N = B.shape[0]
    D = out if out is not None else torch.empty(M, N, device=A.device, dtype=torch.float32)
    call("l2s_gemm_f32", ptr(A), ptr(B), ptr(D), M, N, K, K, 1, K, 1, N, int(accumulate), stream())
    return D

"""RPN proposal layer on the device -- drop-in for pyutils/mask-faster-rcnn/lib/layer_utils/proposal_layer.py:19-68.

Same signature and outputs (blob (N,5) [0,x1,y1,x2,y2], scores (N,1)).  Box decoding + clipping is one kernel
(l2s_proposal_decode), the score ordering is the framework's device sort, and NMS is l2s_nms: bit mask AND greedy scan
on the device, so the only host interaction left is reading the number of survivors (the reference copies the whole
N x N/64 mask to the host and scans it there, nms_cuda.c:38-56).  `proposal_layer_padded` avoids even that and is
CUDA-graph capturable.
"""
import torch

from .. import _lib
from .. import functional as L2F
from .._lib import call, f32c, ptr, stream

# model/config.py:142-148,192-198
CFG = {"TRAIN": dict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7),
       "TEST": dict(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7)}


def decode_proposals(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, num_anchors):
    """bbox_transform_inv + clip_boxes (:41-46) -> (HWA,5) rows [x1,y1,x2,y2,score] (the layout gpu_nms takes)."""
    scores = f32c(rpn_cls_prob[:, :, :, num_anchors:]).reshape(-1)          # positive-class scores, (HWA,)
    deltas = f32c(rpn_bbox_pred).reshape(-1, 4)
    anchors = f32c(anchors).reshape(-1, 4)
    N = scores.numel()
    assert deltas.shape[0] == N and anchors.shape[0] == N, "anchors / deltas / scores disagree on H*W*A"
    info = im_info[0] if getattr(im_info, "ndim", 1) == 2 else im_info
    boxes5 = torch.empty(N, 5, device=scores.device, dtype=torch.float32)
    call("l2s_proposal_decode", ptr(anchors), ptr(deltas), ptr(scores), 1, ptr(boxes5), N, float(info[0]), float(info[1]),
         stream())
    return boxes5


def proposal_layer_padded(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key, _feat_stride, anchors, num_anchors, cfg=None):
    """No host synchronisation: returns (blob (post_nms_topN,5) zero padded, scores (post_nms_topN,1), count (1,) int64)."""
    if isinstance(cfg_key, bytes):
        cfg_key = cfg_key.decode("utf-8")
    c = (cfg or CFG)[cfg_key]
    boxes5 = decode_proposals(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, num_anchors)
    order = torch.sort(boxes5[:, 4], descending=True, stable=True)[1]       # :49
    if c["RPN_PRE_NMS_TOP_N"] > 0:
        order = order[:c["RPN_PRE_NMS_TOP_N"]]
    cand = boxes5[order]
    post = c["RPN_POST_NMS_TOP_N"] if c["RPN_POST_NMS_TOP_N"] > 0 else cand.shape[0]
    keep, num = L2F.nms_sorted(cand, c["RPN_NMS_THRESH"], max_out=post)     # :56 ; positions in score order
    n = min(post, cand.shape[0])
    valid = (torch.arange(n, device=cand.device) < num).unsqueeze(1)
    sel = cand[keep[:n].clamp_(0, max(cand.shape[0] - 1, 0))] * valid
    blob = torch.cat([sel.new_zeros(n, 1), sel[:, :4]], 1)                   # batch index 0 (:65)
    return blob, sel[:, 4:5], torch.minimum(num, num.new_full((1,), n))


def proposal_layer(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key, _feat_stride, anchors, num_anchors, cfg=None):
    """The reference's signature and (variable-size) result; one device->host read of the survivor count."""
    blob, scores, count = proposal_layer_padded(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key, _feat_stride, anchors,
                                                num_anchors, cfg)
    n = int(count)
    return blob[:n], scores[:n]

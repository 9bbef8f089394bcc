"""Proposal target layer on the device -- drop-in for
pyutils/mask-faster-rcnn/lib/layer_utils/proposal_target_layer.py:21-205.

What changes against the reference: overlaps / assignment, box targets and mask targets are device kernels
(l2s_roi_gt_overlaps, l2s_bbox_targets, l2s_mask_crop_resize); the fg / bg sampling draws its randomness from device
priorities instead of numpy.random.choice on the host (pass `rand` for a reproducible / testable draw); the ground-truth
masks are a device uint8 tensor instead of a numpy array, so no per-ROI host loop and no H2D copy of the targets.
The background test is `(ov < HI) & (ov >= LO)`: the reference writes `(a + b) == 2` on ByteTensors (:146), which is the
same thing under PyTorch 0.3 and silently wrong under bool semantics (SURVEY trap T2).
"""
import torch

from .. import functional as L2F
from .._lib import call, f32c, ptr, stream

# model/config.py:49-114
CFG = dict(USE_GT=False, BATCH_SIZE=256, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,
           BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True, BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0),
           BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2), BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), MASK_SIZE=14)


def roi_gt_overlaps(rois, gt_boxes):
    """(max overlap (N,), first arg-max (N,) int64) of rois[:,1:5] against gt_boxes[:,:4] (utils/bbox.pyx)."""
    rois, gt = f32c(rois), f32c(gt_boxes)
    N, G = rois.shape[0], gt.shape[0]
    ov = torch.empty(N, device=rois.device, dtype=torch.float32)
    arg = torch.empty(N, device=rois.device, dtype=torch.int64)
    call("l2s_roi_gt_overlaps", ptr(rois), rois.shape[1], 1, ptr(gt), gt.shape[1], ptr(ov), ptr(arg), N, G, stream())
    return ov, arg


def bbox_regression_targets(rois, gt_boxes, gt_assignment, labels, num_classes, cfg=CFG):
    """_compute_targets + _get_bbox_regression_labels (:83-121) -> (targets (N,4K), inside weights (N,4K))."""
    import ctypes
    rois, gt = f32c(rois), f32c(gt_boxes)
    N = rois.shape[0]
    lab = f32c(labels).reshape(-1)
    tg = torch.empty(N, 4 * num_classes, device=rois.device, dtype=torch.float32)
    iw = torch.empty_like(tg)
    means = cfg["BBOX_NORMALIZE_MEANS"] if cfg["BBOX_NORMALIZE_TARGETS_PRECOMPUTED"] else (0., 0., 0., 0.)
    stds = cfg["BBOX_NORMALIZE_STDS"] if cfg["BBOX_NORMALIZE_TARGETS_PRECOMPUTED"] else (1., 1., 1., 1.)
    f4 = ctypes.c_float * 4
    call("l2s_bbox_targets", ptr(rois), rois.shape[1], 1, ptr(gt), gt.shape[1], ptr(gt_assignment.contiguous()), ptr(lab),
         ptr(tg), ptr(iw), N, num_classes, f4(*means), f4(*stds), f4(*cfg["BBOX_INSIDE_WEIGHTS"]), stream())
    return tg, iw


def _choose(inds, prio, k):
    """k of `inds` without replacement: the k smallest device priorities, in priority order (npr.choice stand-in)."""
    order = torch.sort(prio[inds], stable=True)[1][:k]
    return inds[order]


def _sample_rois(all_rois, all_scores, gt_boxes, gt_masks, fg_rois_per_image, rois_per_image, num_classes, rand, cfg):
    max_overlaps, gt_assignment = roi_gt_overlaps(all_rois, gt_boxes)                    # :137-139
    labels = gt_boxes[gt_assignment, 4]
    fg_inds = (max_overlaps >= cfg["FG_THRESH"]).nonzero().view(-1)                      # :143
    bg_inds = ((max_overlaps < cfg["BG_THRESH_HI"]) & (max_overlaps >= cfg["BG_THRESH_LO"])).nonzero().view(-1)   # :146
    n_all = all_rois.shape[0]
    if rand is None:
        rand = {"fg": torch.rand(n_all, device=all_rois.device), "bg": torch.rand(n_all, device=all_rois.device),
                "replace": torch.rand(int(rois_per_image), device=all_rois.device)}
    n_fg, n_bg = fg_inds.numel(), bg_inds.numel()
    if n_fg > 0 and n_bg > 0:                                                            # :149-154
        fg_rois_per_image = min(fg_rois_per_image, n_fg)
        fg_inds = _choose(fg_inds, rand["fg"], int(fg_rois_per_image))
        bg_need = int(rois_per_image - fg_rois_per_image)
        if n_bg < bg_need:
            bg_inds = bg_inds[(rand["replace"][:bg_need] * n_bg).long().clamp_(max=n_bg - 1)]
        else:
            bg_inds = _choose(bg_inds, rand["bg"], bg_need)
    elif n_fg > 0:                                                                       # :155-158
        need = int(rois_per_image)
        if n_fg < need:
            fg_inds = fg_inds[(rand["replace"][:need] * n_fg).long().clamp_(max=n_fg - 1)]
        else:
            fg_inds = _choose(fg_inds, rand["fg"], need)
        fg_rois_per_image = rois_per_image
    else:                                                                                # :159-168: no fg -> add the gt boxes
        zeros = all_rois.new_zeros(gt_boxes.shape[0], 1)
        all_rois = torch.cat((all_rois, torch.cat((zeros, gt_boxes[:, :-1]), 1)), 0)
        all_scores = torch.cat((all_scores.reshape(-1, 1), zeros), 0)
        rand2 = None if rand is None else {k: (torch.cat([v, v.new_full((gt_boxes.shape[0],), 0.5)]) if k != "replace" else v)
                                           for k, v in rand.items()}
        return _sample_rois(all_rois, all_scores, gt_boxes, gt_masks, fg_rois_per_image, rois_per_image, num_classes,
                            rand2, cfg)
    keep_inds = torch.cat([fg_inds, bg_inds], 0)                                         # :177
    labels = labels[keep_inds].contiguous()
    labels[int(fg_rois_per_image):] = 0
    rois = all_rois[keep_inds].contiguous()
    roi_scores = all_scores.reshape(-1)[keep_inds].contiguous()
    bbox_targets, bbox_inside_weights = bbox_regression_targets(rois, gt_boxes, gt_assignment[keep_inds], labels,
                                                                num_classes, cfg)       # :184-188
    mask_targets = L2F.mask_targets(gt_masks, all_rois[fg_inds], gt_assignment[fg_inds], cfg["MASK_SIZE"])   # :190-201
    return labels, rois, roi_scores, bbox_targets, bbox_inside_weights, mask_targets


def proposal_target_layer(rpn_rois, rpn_scores, gt_boxes, gt_masks, _num_classes, rand=None, cfg=CFG):
    """rpn_rois (N,5) [0,x1,y1,x2,y2] ; rpn_scores (N,) ; gt_boxes (M,5) [x1,y1,x2,y2,cls] ; gt_masks (M,imH,imW) uint8
    DEVICE tensor -> rois, roi_scores, labels (Nkp,1), bbox_targets, bbox_inside_weights, bbox_outside_weights,
    mask_targets (n_fg,14,14)   (the reference's return tuple, :64-81)."""
    all_rois, all_scores = rpn_rois, rpn_scores
    if cfg["USE_GT"]:
        zeros = rpn_rois.new_zeros(gt_boxes.shape[0], 1)
        all_rois = torch.cat((all_rois, torch.cat((zeros, gt_boxes[:, :-1]), 1)), 0)
        all_scores = torch.cat((all_scores.reshape(-1, 1), zeros), 0)
    rois_per_image = cfg["BATCH_SIZE"]
    fg_rois_per_image = int(round(cfg["FG_FRACTION"] * rois_per_image))
    labels, rois, roi_scores, bbox_targets, bbox_inside_weights, mask_targets = _sample_rois(
        all_rois, all_scores, gt_boxes, gt_masks, fg_rois_per_image, rois_per_image, _num_classes, rand, cfg)
    bbox_outside_weights = (bbox_inside_weights > 0).float()
    return (rois.view(-1, 5), roi_scores.view(-1), labels.view(-1, 1), bbox_targets.view(-1, _num_classes * 4),
            bbox_inside_weights.view(-1, _num_classes * 4), bbox_outside_weights, mask_targets)

"""RPN tail and proposal targets on the device against the numpy restatement of
layer_utils/proposal_layer.py and layer_utils/proposal_target_layer.py -- index results bit exact."""
import numpy as np
import pytest
import torch

from conftest import relerr
from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _anchors(H, W, A, g):
    cx = (torch.arange(W).float() * 16 + 8)[None, :, None].expand(H, W, A)
    cy = (torch.arange(H).float() * 16 + 8)[:, None, None].expand(H, W, A)
    s = (32 + 200 * torch.rand(A, generator=g))[None, None, :].expand(H, W, A)
    r = (0.5 + 1.5 * torch.rand(A, generator=g))[None, None, :].expand(H, W, A)
    w, h = s * r.sqrt(), s / r.sqrt()
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 3).reshape(-1, 4).contiguous()


@pytest.mark.parametrize("cfg_key,H,W", [("TEST", 32, 32), ("TRAIN", 38, 63)])
def test_proposal_layer_vs_numpy(cfg_key, H, W):
    from lang2seg_b200.layer_utils.proposal_layer import CFG, decode_proposals, proposal_layer, proposal_layer_padded
    g = torch.Generator().manual_seed(H + W)
    A = 9
    anchors = _anchors(H, W, A, g)
    prob = torch.rand(1, H, W, 2 * A, generator=g)
    deltas = torch.randn(1, H, W, 4 * A, generator=g) * 0.3
    im_info = np.array([[H * 16, W * 16, 1.0]], dtype=np.float32)
    boxes5 = decode_proposals(prob.cuda(), deltas.cuda(), im_info, anchors.cuda(), A).cpu().numpy()
    ref = R.bbox_transform_inv_clip(anchors.numpy(), deltas.reshape(-1, 4).numpy(), H * 16, W * 16)
    # exp() differs by <= 2 ulp between numpy and the device; everything else is the same fp32 arithmetic
    assert np.max(np.abs(boxes5[:, :4] - ref)) <= 1e-3 and relerr(boxes5[:, :4], ref) < 1e-6
    assert np.array_equal(boxes5[:, 4], prob[0, :, :, A:].reshape(-1).numpy())
    c = CFG[cfg_key]
    blob, scores = proposal_layer(prob.cuda(), deltas.cuda(), im_info, cfg_key, 16, anchors.cuda(), A)
    sel = R.proposal_select(boxes5, c["RPN_PRE_NMS_TOP_N"], c["RPN_POST_NMS_TOP_N"], c["RPN_NMS_THRESH"])
    assert blob.shape == (len(sel), 5) and len(sel) > 10
    assert np.array_equal(blob.cpu().numpy()[:, 1:], boxes5[sel, :4])        # same proposals, same order: bit exact
    assert np.array_equal(scores.cpu().numpy()[:, 0], boxes5[sel, 4])
    assert float(blob[:, 0].abs().max()) == 0.0
    pblob, pscores, count = proposal_layer_padded(prob.cuda(), deltas.cuda(), im_info, cfg_key, 16, anchors.cuda(), A)
    assert int(count) == len(sel) and torch.equal(pblob[:len(sel)], blob) and float(pblob[len(sel):].abs().sum()) == 0.0


@pytest.mark.parametrize("case", ["mixed", "few_bg", "only_fg"])
def test_proposal_target_layer_vs_numpy(case):
    from lang2seg_b200.layer_utils.proposal_target_layer import CFG, proposal_target_layer
    g = torch.Generator().manual_seed(len(case))
    imH, imW, G, K = 320, 480, 5, 81
    gt = torch.zeros(G, 5)
    gt[:, 0] = torch.rand(G, generator=g) * 300
    gt[:, 1] = torch.rand(G, generator=g) * 200
    gt[:, 2] = gt[:, 0] + 40 + torch.rand(G, generator=g) * 120
    gt[:, 3] = gt[:, 1] + 40 + torch.rand(G, generator=g) * 100
    gt[:, 2].clamp_(max=imW - 1); gt[:, 3].clamp_(max=imH - 1)
    gt[:, 4] = torch.randint(1, K, (G,), generator=g).float()
    masks = (torch.rand(G, imH, imW, generator=g) < 0.5).to(torch.uint8)
    n = {"mixed": 600, "few_bg": 300, "only_fg": 300}[case]
    # candidates: jittered copies of the gt boxes (foreground-ish) and random boxes (background-ish)
    src = gt[torch.randint(0, G, (n,), generator=g), :4]
    jit = {"mixed": 25.0, "few_bg": 6.0, "only_fg": 2.0}[case]
    boxes = src + torch.randn(n, 4, generator=g) * jit
    n_rnd = {"mixed": n // 2, "few_bg": 60, "only_fg": 0}[case]
    if n_rnd:
        rnd = torch.rand(n_rnd, 4, generator=g) * torch.tensor([imW * 0.6, imH * 0.6, imW * 0.4, imH * 0.4])
        boxes[:n_rnd] = torch.stack([rnd[:, 0], rnd[:, 1], rnd[:, 0] + rnd[:, 2] + 8, rnd[:, 1] + rnd[:, 3] + 8], 1)
    boxes[:, 0::2] = boxes[:, 0::2].clamp(0, imW - 1); boxes[:, 1::2] = boxes[:, 1::2].clamp(0, imH - 1)
    boxes[:, 2] = torch.maximum(boxes[:, 2], boxes[:, 0]); boxes[:, 3] = torch.maximum(boxes[:, 3], boxes[:, 1])
    rois = torch.cat([torch.zeros(n, 1), boxes], 1)
    scores = torch.rand(n, generator=g)
    rand = {"fg": torch.rand(n, generator=g), "bg": torch.rand(n, generator=g), "replace": torch.rand(256, generator=g)}
    ref = R.sample_rois_np(rois.numpy(), scores.numpy(), gt.numpy(), masks.numpy(), {k: v.numpy() for k, v in rand.items()}, K, CFG)
    if case == "few_bg":
        assert 0 < (ref["max_overlaps"] < 0.5).sum() < 192      # sampling WITH replacement is exercised
    if case == "only_fg":
        assert (ref["max_overlaps"] < 0.5).sum() == 0
    out = proposal_target_layer(rois.cuda(), scores.cuda(), gt.cuda(), masks.cuda(), K, rand={k: v.cuda() for k, v in rand.items()})
    o_rois, o_scores, o_labels, o_tg, o_iw, o_ow, o_mt = [t.cpu() for t in out]
    assert np.array_equal(o_rois.numpy(), ref["rois"])                     # same sampled ROIs in the same order
    assert np.array_equal(o_labels.numpy()[:, 0], ref["labels"])
    assert np.array_equal(o_scores.numpy(), ref["roi_scores"])
    assert np.array_equal(o_iw.numpy(), ref["bbox_inside_weights"]) and np.array_equal(o_ow.numpy(), ref["bbox_outside_weights"])
    assert np.max(np.abs(o_tg.numpy() - ref["bbox_targets"])) < 1e-5       # log() differs by <= 2 ulp
    assert o_mt.shape == ref["mask_targets"].shape and np.array_equal(o_mt.numpy(), ref["mask_targets"])


def test_proposal_target_layer_without_foreground_appends_gt():
    """:159-168 -- no candidate reaches FG_THRESH: the ground-truth boxes join the candidates and become the foreground."""
    from lang2seg_b200.layer_utils.proposal_target_layer import proposal_target_layer
    gt = torch.tensor([[100., 100., 200., 220., 7.], [300., 50., 420., 150., 3.]])
    masks = torch.ones(2, 320, 480, dtype=torch.uint8)
    g = torch.Generator().manual_seed(4)
    xy = torch.rand(200, 2, generator=g) * 40
    rois = torch.cat([torch.zeros(200, 1), xy, xy + 20], 1)                # far away from both gt boxes
    out = proposal_target_layer(rois.cuda(), torch.rand(200, generator=g).cuda(), gt.cuda(), masks.cuda(), 81)
    labels = out[2].cpu()[:, 0]
    assert out[0].shape == (256, 5) and sorted(labels[labels > 0].tolist()) == [3.0, 7.0]
    assert out[6].shape == (2, 14, 14) and float(out[6].min()) == 1.0

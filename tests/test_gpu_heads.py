"""Box head (row a9, Network._region_classification) and the hot-path loss sum (row a16) vs the oracle."""
import pytest
import torch

from conftest import relerr
from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 1e-4


@pytest.mark.parametrize("N", [3, 100, 768])
def test_region_classification_vs_oracle(N):
    """7x7 mean -> stacked Linears (skinny fp32 GEMM for N < 512, tcgen05 bf16x3 above) -> softmax / argmax."""
    from lang2seg_b200.nets.network import HotPathNet
    torch.manual_seed(1)
    net = HotPathNet().cuda().eval()
    with torch.no_grad():      # the reference's N(0, 0.01 / 0.001) init gives near-uniform softmax: widen it for the test
        net.cls_score_net.weight.mul_(20.)
        net.bbox_pred_net.weight.mul_(50.)
        net.cls_score_net.bias.normal_(0, 0.1)
        net.bbox_pred_net.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(N)
    x = torch.relu(torch.randn(N, 2048, 7, 7, generator=g))
    Gs, Gb = torch.randn(N, 81, generator=g), torch.randn(N, 324, generator=g)
    xc = x.cuda().requires_grad_(True)
    cls_prob, bbox_pred = net._region_classification(xc)
    cls_score = net._predictions["cls_score"]
    ((cls_score * Gs.cuda()).sum() + (bbox_pred * Gb.cuda()).sum()).backward()

    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()
         if k.startswith(("cls_score_net", "bbox_pred_net"))}
    xo = x.clone().requires_grad_(True)
    so, _, po, bo = R.region_classification(xo, p["cls_score_net.weight"], p["cls_score_net.bias"],
                                         p["bbox_pred_net.weight"], p["bbox_pred_net.bias"])
    ((so * Gs).sum() + (bo * Gb).sum()).backward()
    assert relerr(cls_score, so) < TOL and relerr(cls_prob, po) < TOL and relerr(bbox_pred, bo) < TOL
    # argmax: identical wherever the two best scores are not within rounding of each other
    top2 = so.detach().topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * so.detach().abs().max()
    assert torch.equal(net._predictions["cls_pred"].cpu()[clear], so.detach().argmax(1)[clear])
    assert relerr(xc.grad, xo.grad) < TOL
    assert relerr(net.cls_score_net.weight.grad, p["cls_score_net.weight"].grad) < TOL
    assert relerr(net.bbox_pred_net.weight.grad, p["bbox_pred_net.weight"].grad) < TOL
    assert relerr(net.cls_score_net.bias.grad, p["cls_score_net.bias"].grad) < TOL
    assert relerr(net.bbox_pred_net.bias.grad, p["bbox_pred_net.bias"].grad) < TOL


def test_spatial_mean_shapes_and_softmax_ties():
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(2)
    for shape in [(5, 7, 7, 7), (33, 40, 3, 5), (1, 1, 1, 1)]:
        x = torch.randn(*shape, generator=g)
        xc = x.cuda().requires_grad_(True)
        out = F.spatial_mean(xc)
        ref = x.mean(3).mean(2)
        assert relerr(out, ref) < 1e-6
        G = torch.randn(out.shape, generator=g)
        (gx,) = torch.autograd.grad((out * G.cuda()).sum(), xc)
        assert relerr(gx, (G / (shape[2] * shape[3]))[:, :, None, None].expand(shape)) < 1e-6
    s = torch.tensor([[1., 3., 3., 0.], [2., 2., 2., 2.], [-1., -5., -1., -7.]]).cuda()
    prob, pred = F.softmax_argmax(s)
    assert pred.tolist() == [1, 0, 0]                       # first maximum, as torch.max
    assert relerr(prob, torch.softmax(s, 1)) < 1e-6


def test_add_hot_path_losses_matches_reference_sum():
    """Row a16: the weighted loss sum of network_cycle_response.py:449 from the tensors the forward methods left behind,
    E = 1 (the reference's case), cap_loss_weight != 1."""
    import torch.nn.functional as TF
    from lang2seg_b200.nets.network import HotPathNet
    torch.manual_seed(3)
    net = HotPathNet(dict(cap_loss_weight=0.3)).cuda().eval()
    g = torch.Generator().manual_seed(9)
    N, nfg, H, W, L = 24, 6, 32, 32, 10
    X = torch.relu(torch.randn(1, 1024, H, W, generator=g))
    labels, lens = R.synth_labels(g, 1, L, 1999)
    cap, msk = R.caption_targets(labels, lens, L)
    fc7 = torch.relu(torch.randn(N, 2048, 7, 7, generator=g))
    rtgt = (torch.rand(1, H, W, generator=g) < 0.3).float()
    fc, att = torch.randn(1, 4096, generator=g), torch.relu(torch.randn(1, 14, 14, 4096, generator=g))
    tg = {"labels": torch.randint(0, 81, (N,), generator=g), "bbox_targets": torch.randn(N, 324, generator=g) * 0.1,
          "bbox_inside_weights": (torch.rand(N, 324, generator=g) < 0.2).float(),
          "mask_targets": (torch.rand(nfg, 14, 14, generator=g) < 0.5).float()}
    tg["labels"][:nfg] = torch.randint(1, 81, (nfg,), generator=g)
    tg["bbox_outside_weights"] = tg["bbox_inside_weights"].clone()

    net._dynamic_filter(X.cuda(), labels.cuda(), expr2img=torch.zeros(1, dtype=torch.int32), resp_target=rtgt.cuda())
    net._region_classification(fc7.cuda())
    net._proposal_targets.update({k: v.cuda() for k, v in tg.items()})
    net._mask_prediction(fc7[:nfg].cuda(), tg["labels"][:nfg].cuda(), tg["mask_targets"].cuda())
    total = net._add_hot_path_losses(cap.cuda(), msk.cuda(), fc.cuda(), att.cuda(), rpn_cross_entropy=0.25, rpn_loss_box=0.125)

    p = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    enc = {k[len("rnn_encoder."):]: v for k, v in p.items() if k.startswith("rnn_encoder.")}
    _, hidden, _ = R.rnn_encoder(labels, enc)
    filt, fuse = R.filter_generator(hidden, [p["dynamic_fc_%d.weight" % k] for k in range(7)],
                                    [p["dynamic_fc_%d.bias" % k] for k in range(7)], p["response_fc.weight"], p["response_fc.bias"])
    r, _ = R.dynamic_filter(X, filt, fuse, [0])
    so, _, _, bo = R.region_classification(fc7, p["cls_score_net.weight"], p["cls_score_net.bias"],
                                           p["bbox_pred_net.weight"], p["bbox_pred_net.bias"])
    s, _ = R.mask_head(fc7[:nfg], p["mask_up_sampling.weight"], p["mask_up_sampling.bias"], p["mask_pred_net.weight"],
                       p["mask_pred_net.bias"])
    capp = {k[len("caption_model."):]: v for k, v in p.items() if k.startswith("caption_model.")}
    diff = tg["bbox_inside_weights"] * (bo - tg["bbox_targets"])
    sl1 = torch.where(diff.abs() < 1, 0.5 * diff ** 2, diff.abs() - 0.5)
    ref = (TF.cross_entropy(so, tg["labels"]) + (tg["bbox_outside_weights"] * sl1).sum(1).mean() + 0.25 + 0.125
           + R.mask_loss(s, tg["labels"][:nfg], tg["mask_targets"]) + R.response_loss(r, rtgt).sum()
           + 0.3 * R.caption_loss(fc, att, cap, msk, capp))
    assert relerr(total, ref) < TOL
    for k in ("cross_entropy", "loss_box", "loss_mask", "loss_response", "loss_caption", "total_loss"):
        assert k in net._losses

"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN MODULES (under oracle/shim.py).

Run in the development container only (needs /root/reference):

    python -m oracle.make_golden

Each fixture holds seeded inputs, the parameters of the reference modules involved and
the reference's outputs / autograd gradients on them.  Sizes are kept small (reduced
channel counts / hidden sizes through the reference's own constructor options) so the
fixtures stay a few hundred KB; the arithmetic path through the reference code is the
same as at full size.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_OPT = dict(vocab_size=40, word_embedding_size=12, word_vec_size=12, rnn_hidden_size=8,
                 C4_feat_dim=20, input_encoding_size=16, rnn_size=16, att_hid_size=16,
                 fc_feat_size=24, att_feat_size=24, seq_length=6)


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _save(name, d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **_np(d))
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def golden_lang_encoder():
    enc = shim.reference_rnn_encoder(SMALL_OPT, seed=11)
    g = torch.Generator().manual_seed(12)
    labels = torch.randint(1, 40, (4, 7), generator=g)
    lens = torch.tensor([7, 3, 5, 1])
    labels = labels * (torch.arange(7)[None] < lens[:, None])
    with torch.no_grad():
        out, hid, emb = enc(labels)
    d = {"labels": labels, "output": out, "hidden": hid, "embedded": emb}
    d.update({"p." + k: v for k, v in enc.state_dict().items()})
    _save("lang_encoder.npz", d)


def golden_dynfilter():
    """Reference Network._predict lines 503-570 with C4_feat_dim=20 and the tail stubbed (D6/D7)."""
    net = shim.build_reference_net(SMALL_OPT, num_layers=50, seed=21)
    # stub everything after the gating so that only :503-570 run with the small channel count
    net._head_to_tail = lambda pool5: pool5
    net._region_classification = lambda x: (None, None)
    net._mask_prediction = lambda x: None
    net._crop_pool_layer = lambda bottom, rois: bottom[:, :, :1, :1]
    g = torch.Generator().manual_seed(22)
    res = {}
    for tag, (H, W) in {"a": (9, 13), "b": (8, 8)}.items():
        X = torch.relu(torch.randn(1, 20, H, W, generator=g)).requires_grad_(True)
        labels = torch.randint(1, 40, (1, 5), generator=g)
        rois = torch.tensor([[0, 0., 0., 31., 31.]])
        G = torch.randn(1, 20, H, W, generator=g)
        tgt = (torch.rand(H, W, generator=g) < 0.3).float()
        net.zero_grad()
        Y = shim.run_predict(net, X, labels, rois, mode="TEST")[0]
        r = net._predictions["response"]
        # loss = <Y,G> + response BCE (network_cycle_response.py:415-422 op on the stored response)
        loss = (Y * G).sum() + F.binary_cross_entropy_with_logits(r.squeeze(1).squeeze(0), tgt)
        loss.backward()
        _, hidden, _ = net.rnn_encoder(labels)
        res.update({tag + ".X": X, tag + ".labels": labels, tag + ".G": G, tag + ".tgt": tgt,
                    tag + ".hidden": hidden, tag + ".response": r, tag + ".Y": Y, tag + ".dX": X.grad,
                    tag + ".d_dyn3_w": net.dynamic_fc_3.weight.grad.clone(),
                    tag + ".d_dyn0_b": net.dynamic_fc_0.bias.grad.clone(),
                    tag + ".d_resp_w": net.response_fc.weight.grad.clone()})
    for k in range(7):
        res["p.dynamic_fc_%d.weight" % k] = getattr(net, "dynamic_fc_%d" % k).weight
        res["p.dynamic_fc_%d.bias" % k] = getattr(net, "dynamic_fc_%d" % k).bias
    res["p.response_fc.weight"] = net.response_fc.weight
    res["p.response_fc.bias"] = net.response_fc.bias
    _save("dynfilter.npz", res)


def golden_crop():
    """Reference Network._crop_pool_layer / _crop_pool_layer_align (:107-182) + autograd backward."""
    install_net = types.SimpleNamespace()
    shim.install()
    from nets.network_cycle_response import Network
    g = torch.Generator().manual_seed(31)
    H, W = 9, 13
    bottom = torch.randn(1, 6, H, W, generator=g).requires_grad_(True)
    rois = torch.tensor([
        [0, 10.0, 20.0, 150.0, 120.0],
        [0, 0.0, 0.0, 207.0, 143.0],       # beyond the right / bottom edge -> zero padding
        [0, 16.0, 32.0, 96.0, 64.0],       # exact integer feature coordinates
        [0, 50.0, 40.0, 50.0, 40.0],       # degenerate point box
        [0, 120.0, 100.0, 60.0, 30.0],     # inverted box (x2<x1, y2<y1)
        [0, 3.3, 7.7, 191.9, 127.2],
        [0, 100.5, 60.25, 130.75, 90.5],
        [0, 192.0, 128.0, 192.0, 128.0],   # exactly the last pixel (W-1)*16
    ])
    res = {"bottom": bottom, "rois": rois, "im_info": np.array([[144.0, 208.0, 1.0]], dtype=np.float32)}
    for tag, fn in {
        "p7": lambda: Network._crop_pool_layer(install_net, bottom, rois, False),
        "p14max": lambda: Network._crop_pool_layer(install_net, bottom, rois, True),
        "align7": lambda: Network._crop_pool_layer_align(install_net, bottom, rois, res["im_info"], False),
        "align14max": lambda: Network._crop_pool_layer_align(install_net, bottom, rois, res["im_info"], True),
    }.items():
        out = fn()
        G = torch.randn(out.shape, generator=g)
        (gb,) = torch.autograd.grad((out * G).sum(), bottom)
        res.update({tag + ".out": out, tag + ".G": G, tag + ".dbottom": gb})
    _save("crop.npz", res)


def golden_mask_head():
    """Reference Network._mask_prediction (:292-307) on small layers of the same module types.

    The mask loss (:404-413) lives inside the monolithic _add_losses; its three torch ops
    (gather, squeeze, BCE-with-logits) are applied here in the same order.
    """
    shim.install()
    from nets.network_cycle_response import Network
    torch.manual_seed(41)
    ns = types.SimpleNamespace(mask_up_sampling=nn.ConvTranspose2d(40, 24, 2, 2),
                               mask_pred_net=nn.Conv2d(24, 5, kernel_size=1, stride=1), _predictions={})
    g = torch.Generator().manual_seed(42)
    x = torch.relu(torch.randn(3, 40, 7, 7, generator=g)).requires_grad_(True)
    labels = torch.tensor([1, 4, 2])
    tgt = (torch.rand(3, 14, 14, generator=g) < 0.5).float()
    prob = Network._mask_prediction(ns, x)
    score = ns._predictions["mask_score"]
    idx = labels.view(3, 1, 1, 1).expand(3, 1, 14, 14)
    loss = F.binary_cross_entropy_with_logits(torch.gather(score, 1, idx).squeeze(1), tgt)
    loss.backward()
    _save("mask_head.npz", {
        "x": x, "labels": labels, "tgt": tgt, "score": score, "prob": prob, "loss": loss, "dx": x.grad,
        "up_w": ns.mask_up_sampling.weight, "up_b": ns.mask_up_sampling.bias,
        "pred_w": ns.mask_pred_net.weight, "pred_b": ns.mask_pred_net.bias,
        "d_up_w": ns.mask_up_sampling.weight.grad, "d_up_b": ns.mask_up_sampling.bias.grad,
        "d_pred_w": ns.mask_pred_net.weight.grad, "d_pred_b": ns.mask_pred_net.bias.grad})


def golden_att2in2():
    """Reference Att2in2Model.forward + LanguageModelCriterion, eval mode, with the early break."""
    model = shim.reference_caption_model(SMALL_OPT, seed=51)
    import misc.utils as mutils
    g = torch.Generator().manual_seed(52)
    B, L = 3, 6
    labels = torch.randint(1, 40, (B, L), generator=g)
    lens = torch.tensor([4, 2, 3])                       # max 4 < L: T = 5 by the break rule
    labels = labels * (torch.arange(L)[None] < lens[:, None])
    cap = torch.zeros(B, L + 2, dtype=torch.long)
    cap[:, 1:L + 1] = labels
    msk = (torch.arange(L + 2)[None] < (lens[:, None] + 2)).float()
    fc = torch.randn(B, 24, generator=g)
    att = torch.relu(torch.randn(B, 3, 4, 24, generator=g)).requires_grad_(True)
    logp = model(fc, att, cap)
    loss = mutils.LanguageModelCriterion()(logp, cap[:, 1:], msk[:, 1:])
    loss.backward()
    # one isolated attention step (AttModel.py:406-423)
    h = torch.randn(B, 16, generator=g)
    af = torch.randn(B, 12, 16, generator=g)
    pf = torch.randn(B, 12, 16, generator=g)
    with torch.no_grad():
        att_res = model.core.attention(h, af, pf)
    d = {"fc": fc, "att": att, "cap": cap, "msk": msk, "logp": logp, "loss": loss, "d_att": att.grad,
         "step.h": h, "step.att_feats": af, "step.p_att": pf, "step.att_res": att_res}
    for k, v in model.state_dict().items():
        d["p." + k] = v
    for k, v in model.named_parameters():
        d["g." + k] = v.grad
    _save("att2in2.npz", d)


def golden_caption_features():
    """network_cycle_response.py:428-439 ops on small maps (mean, adaptive_avg_pool2d, permute, cat)."""
    g = torch.Generator().manual_seed(61)
    fb = torch.randn(1, 10, 19, 32, generator=g)
    fa = torch.randn(1, 10, 19, 32, generator=g)
    fc_b = fb.mean(3).mean(2)
    att_b = F.adaptive_avg_pool2d(fb, [14, 14]).permute(0, 2, 3, 1).contiguous()
    fc_a = fa.mean(3).mean(2)
    att_a = F.adaptive_avg_pool2d(fa, [14, 14]).permute(0, 2, 3, 1).contiguous()
    _save("caption_features.npz", {"fb": fb, "fa": fa, "fc": torch.cat((fc_b, fc_a), 1),
                                   "att": torch.cat((att_b, att_a), 3)})


def golden_imresize():
    """scipy.misc.imresize(...,'nearest') stand-in (D4) on a {0,1} uint8 mask, as at :418."""
    shim.install()
    import scipy.misc
    g = np.random.RandomState(71)
    m = (g.rand(37, 53) < 0.4).astype(np.uint8)
    _save("imresize.npz", {"mask": m, "r9x13": scipy.misc.imresize(m, (9, 13), interp="nearest"),
                           "r32x32": scipy.misc.imresize(m, (32, 32), interp="nearest")})


def main():
    golden_lang_encoder()
    golden_dynfilter()
    golden_crop()
    golden_mask_head()
    golden_att2in2()
    golden_caption_features()
    golden_imresize()


if __name__ == "__main__":
    main()

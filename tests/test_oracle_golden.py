"""The oracle (oracle/restate.py) against the fixtures generated from the reference's own modules."""
import numpy as np
import torch
import torch.nn.functional as F

from conftest import relerr
from oracle import restate as R

TOL = 2e-6   # fp32 CPU vs fp32 CPU, different op order


def _params(d, prefix="p."):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def test_lang_encoder(golden):
    d = golden("lang_encoder.npz")
    out, hid, emb = R.rnn_encoder(d["labels"], _params(d))
    assert relerr(out, d["output"]) < TOL
    assert relerr(hid, d["hidden"]) < TOL
    assert relerr(emb, d["embedded"]) < TOL


def test_lang_encoder_packed_equals_loops(golden):
    """bench.py's CPU baseline runs the encoder through torch's packed nn.LSTM (the reference's own calls): same
    outputs and gradients as the per-token restatement, and as the reference's golden output."""
    import torch
    d = golden("lang_encoder.npz")
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in _params(d).items()}
    out, hid, emb = R.rnn_encoder_packed(d["labels"], p)
    assert relerr(out, d["output"]) < TOL and relerr(hid, d["hidden"]) < TOL and relerr(emb, d["embedded"]) < TOL
    g1 = torch.autograd.grad(hid.pow(2).sum() + out.sum(), [p["rnn.weight_hh_l0"], p["rnn.weight_ih_l0_reverse"]])
    out2, hid2, _ = R.rnn_encoder(d["labels"], p)
    g2 = torch.autograd.grad(hid2.pow(2).sum() + out2.sum(), [p["rnn.weight_hh_l0"], p["rnn.weight_ih_l0_reverse"]])
    for a, b in zip(g1, g2):
        assert relerr(a, b) < 1e-5


def test_partition_bounds():
    # int() of true division
    assert R.partition_bounds(38, 63) == (19, 9, 28, 31, 15, 47)
    assert R.partition_bounds(37, 62) == (18, 9, 27, 31, 15, 46)
    assert R.partition_bounds(9, 13) == (4, 2, 6, 6, 3, 9)


def test_dynfilter_forward_backward(golden):
    d = golden("dynfilter.npz")
    p = _params(d)
    dyn_w = [p["dynamic_fc_%d.weight" % k].clone().requires_grad_(True) for k in range(7)]
    dyn_b = [p["dynamic_fc_%d.bias" % k].clone().requires_grad_(True) for k in range(7)]
    rw = p["response_fc.weight"].clone().requires_grad_(True)
    rb = p["response_fc.bias"]
    for tag in ("a", "b"):
        for fn in (R.dynamic_filter, R.dynamic_filter_closed_form):
            X = d[tag + ".X"].clone().requires_grad_(True)
            f, w = R.filter_generator(d[tag + ".hidden"], dyn_w, dyn_b, rw, rb)
            r, Y = fn(X, f, w)
            assert relerr(r, d[tag + ".response"]) < TOL
            assert relerr(Y, d[tag + ".Y"]) < TOL
            loss = (Y * d[tag + ".G"]).sum() + R.response_loss(r, d[tag + ".tgt"][None]).sum()
            gX, g3, g0b, grw = torch.autograd.grad(loss, [X, dyn_w[3], dyn_b[0], rw])
            assert relerr(gX, d[tag + ".dX"]) < 5e-6
            assert relerr(g3, d[tag + ".d_dyn3_w"]) < 5e-6
            assert relerr(g0b, d[tag + ".d_dyn0_b"]) < 5e-6
            assert relerr(grw, d[tag + ".d_resp_w"]) < 5e-6


def test_crop_pool(golden):
    d = golden("crop.npz")
    imhw = (float(d["im_info"][0, 0]), float(d["im_info"][0, 1]))
    cases = {"p7": dict(max_pool=False), "p14max": dict(max_pool=True),
             "align7": dict(max_pool=False, align_im_hw=imhw), "align14max": dict(max_pool=True, align_im_hw=imhw)}
    for tag, kw in cases.items():
        for fn, tol in ((R.crop_pool, 2e-6), (R.crop_pool_closed_form, 3e-5)):
            b = d["bottom"].clone().requires_grad_(True)
            out = fn(b, d["rois"], **kw)
            assert out.shape == d[tag + ".out"].shape
            assert relerr(out, d[tag + ".out"]) < tol, (tag, fn.__name__)
            (gb,) = torch.autograd.grad((out * d[tag + ".G"]).sum(), b)
            assert relerr(gb, d[tag + ".dbottom"]) < tol * 3, (tag, fn.__name__)


def test_mask_head(golden):
    d = golden("mask_head.npz")
    x = d["x"].clone().requires_grad_(True)
    ws = [d[k].clone().requires_grad_(True) for k in ("up_w", "up_b", "pred_w", "pred_b")]
    s, pr = R.mask_head(x, *ws)
    assert relerr(s, d["score"]) < TOL and relerr(pr, d["prob"]) < TOL
    loss = R.mask_loss(s, d["labels"], d["tgt"])
    assert relerr(loss, d["loss"]) < TOL
    gs = torch.autograd.grad(loss, [x] + ws)
    for g, k in zip(gs, ("dx", "d_up_w", "d_up_b", "d_pred_w", "d_pred_b")):
        assert relerr(g, d[k]) < 5e-6, k


def test_att2in2(golden):
    d = golden("att2in2.npz")
    p = {k: v.clone().requires_grad_(True) for k, v in _params(d).items()}
    att = d["att"].clone().requires_grad_(True)
    assert R.decode_steps(d["cap"]) == d["logp"].shape[1] == 5
    logp = R.att2in2_forward(d["fc"], att, d["cap"], p)
    assert relerr(logp, d["logp"]) < TOL
    loss = R.lm_criterion(logp, d["cap"][:, 1:], d["msk"][:, 1:])
    assert relerr(loss, d["loss"]) < TOL
    loss.backward()
    assert relerr(att.grad, d["d_att"]) < 1e-5
    for k, v in p.items():
        gk = "g." + k
        if k.endswith("alpha_net.bias"):
            continue        # softmax is shift invariant: this gradient is exactly 0 up to rounding noise
        if gk in d and d[gk].numel() and v.grad is not None:
            assert relerr(v.grad, d[gk]) < 1e-5, k
    res, w = R.attention_step(d["step.h"], d["step.att_feats"], d["step.p_att"],
                              p["core.attention.h2att.weight"], p["core.attention.h2att.bias"],
                              p["core.attention.alpha_net.weight"], p["core.attention.alpha_net.bias"])
    assert relerr(res, d["step.att_res"]) < TOL
    assert abs(float(w.sum()) - w.shape[0]) < 1e-5


def test_caption_features(golden):
    d = golden("caption_features.npz")
    fc, att = R.caption_features(d["fb"], d["fa"])
    assert relerr(fc, d["fc"]) < TOL and relerr(att, d["att"]) < TOL
    # integer bin contract of adaptive_avg_pool2d, checked by explicit bins
    hb, wb = R.adaptive_bins(19, 14), R.adaptive_bins(32, 14)
    x = d["fb"][0, 3]
    manual = torch.stack([torch.stack([x[h0:h1, w0:w1].mean() for (w0, w1) in wb]) for (h0, h1) in hb])
    assert relerr(manual, d["att"][0, :, :, 3]) < TOL


def test_imresize_nearest(golden):
    d = golden("imresize.npz")
    m = d["mask"].numpy()
    assert np.array_equal(R.nearest_resize_mask(m, 9, 13), d["r9x13"].numpy().astype(np.float32))
    assert np.array_equal(R.nearest_resize_mask(m, 32, 32), d["r32x32"].numpy().astype(np.float32))


def test_roi_max_pool_properties():
    g = torch.Generator().manual_seed(5)
    f = torch.randn(2, 3, 12, 17, generator=g)
    rois = torch.tensor([[0, 0., 0., 271., 191.], [1, 33., 18., 120., 99.], [0, 200., 150., 90., 60.],
                         [1, 500., 500., 600., 600.], [0, 8., 8., 8., 8.], [1, 24., 40., 25., 200.]])
    out, arg = R.roi_max_pool(f, rois)
    flat = f.reshape(2, -1)
    for n in range(rois.shape[0]):
        b = int(rois[n, 0])
        m = arg[n] >= 0
        assert torch.equal(out[n][m], flat[b][arg[n][m].long()])
        assert torch.all(out[n][~m] == 0)
    # ROI 3 lies outside the map entirely: all bins empty
    assert torch.all(arg[3] == -1)
    # whole-map ROI with 1x1 pooling is the global max
    o1, a1 = R.roi_max_pool(f, rois[:1], 1, 1)
    assert torch.equal(o1[0, :, 0, 0], f[0].flatten(1).max(1)[0])
    gb = R.roi_max_pool_backward(torch.ones_like(out), rois, arg, f.shape)
    assert float(gb.sum()) == float((arg >= 0).sum())


def test_c_oracle_matches_numpy_oracle(golden):
    from oracle import clib
    g = torch.Generator().manual_seed(6)
    f = torch.randn(2, 4, 11, 15, generator=g)
    rois = torch.cat([R.synth_rois(g, 9, 11 * 16, 15 * 16, 0), R.synth_rois(g, 7, 11 * 16, 15 * 16, 1)])
    out, arg = R.roi_max_pool(f, rois)
    o2, a2 = clib.roi_maxpool_fwd(f.numpy(), rois.numpy())
    assert np.array_equal(out.numpy(), o2) and np.array_equal(arg.numpy(), a2)
    top = torch.randn(out.shape, generator=g)
    gb = R.roi_max_pool_backward(top, rois, arg, f.shape)
    assert relerr(clib.roi_maxpool_bwd(top.numpy(), rois.numpy(), a2, tuple(f.shape)), gb) < 1e-6
    d = golden("crop.npz")
    imhw = (float(d["im_info"][0, 0]), float(d["im_info"][0, 1]))
    b, r = d["bottom"].numpy(), d["rois"].numpy()
    assert relerr(clib.crop_resize_fwd(b, r, 7), d["p7.out"]) < 3e-5
    assert relerr(clib.crop_resize_fwd(b, r, 14, True), d["p14max.out"]) < 3e-5
    assert relerr(clib.crop_resize_fwd(b, r, 7, False, imhw), d["align7.out"]) < 3e-5
    assert relerr(clib.crop_resize_fwd(b, r, 14, True, imhw), d["align14max.out"]) < 3e-5


def test_nms_oracle_vs_brute_force():
    """The vectorised fp32 restatement of gpu_nms (nms_cuda.c:44-56 + devIoU nms_kernel.cu:15-24) against a scalar loop."""
    rs = np.random.RandomState(0)
    n = 300
    x1, y1 = rs.uniform(0, 100, n), rs.uniform(0, 100, n)
    b = np.stack([x1, y1, x1 + rs.uniform(5, 60, n), y1 + rs.uniform(5, 60, n), np.linspace(1, 0, n)], 1).astype(np.float32)
    one, zero = np.float32(1), np.float32(0)

    def iou(a, c):
        w = max(np.float32(min(a[2], c[2]) - max(a[0], c[0]) + one), zero)
        h = max(np.float32(min(a[3], c[3]) - max(a[1], c[1]) + one), zero)
        inter = np.float32(w * h)
        sa = np.float32((a[2] - a[0] + one) * (a[3] - a[1] + one))
        sb = np.float32((c[2] - c[0] + one) * (c[3] - c[1] + one))
        return np.float32(inter / np.float32(np.float32(sa + sb) - inter))

    for thresh in (0.3, 0.7):
        removed, keep = [False] * n, []
        for i in range(n):
            if removed[i]:
                continue
            keep.append(i)
            for j in range(i + 1, n):
                if iou(b[i], b[j]) > np.float32(thresh):
                    removed[j] = True
        assert list(R.nms_sorted(b, thresh)) == keep


def test_mask_targets_oracle(golden):
    """proposal_target_layer.py:193-201: python slicing + imresize 'nearest' (pinned by the imresize golden)."""
    d = golden("imresize.npz")
    m = d["mask"].numpy().astype(np.uint8)[None]
    H, W = m.shape[1:]
    rois = np.array([[0, 0, 0, W - 1, H - 1], [0, 3.7, 2.2, 20.9, 30.1], [0, 5.5, 5.5, 5.6, 5.6]], np.float32)
    t = R.mask_targets(m, rois, [0, 0, 0], 14)
    assert np.array_equal(t[0], R.nearest_resize_mask(m[0], 14, 14))
    assert np.array_equal(t[1], R.nearest_resize_mask(m[0, 2:31, 3:21], 14, 14))
    assert np.array_equal(t[2], np.full((14, 14), float(m[0, 5, 5]), np.float32))
    # the integer form used by the kernel equals the float form of the oracle
    for src in range(1, 70):
        for dst in (7, 14, 32):
            assert R.nearest_resize_index(dst, src) == [min(((2 * i + 1) * src) // (2 * dst), src - 1) for i in range(dst)]

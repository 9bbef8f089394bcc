"""Pin oracle/restate.py against the LIVE reference modules (development container only).

Skipped wherever /root/reference is absent (e.g. the GPU box); tests/golden carries the
same evidence there.
"""
import pytest
import torch

from conftest import relerr
from oracle import restate as R
from oracle import shim

pytestmark = pytest.mark.skipif(not shim.available(), reason="reference tree not mounted")


def test_att2in2_full_size_vs_reference():
    model = shim.reference_caption_model(seed=3)
    p = dict(model.state_dict())
    g = torch.Generator().manual_seed(4)
    labels, lens = R.synth_labels(g, 2, 10, 1999)
    cap, msk = R.caption_targets(labels, lens, 10)
    fc = torch.randn(2, 4096, generator=g)
    att = torch.relu(torch.randn(2, 14, 14, 4096, generator=g))
    with torch.no_grad():
        ref = model(fc, att, cap)
        mine = R.att2in2_forward(fc, att, cap, p)
    assert ref.shape == mine.shape == (2, 11, 2000)
    assert relerr(mine, ref) < 1e-5


def test_lang_encoder_full_size_vs_reference():
    enc = shim.reference_rnn_encoder(seed=5)
    g = torch.Generator().manual_seed(6)
    labels, _ = R.synth_labels(g, 3, 10, 1999)
    with torch.no_grad():
        out, hid, emb = enc(labels)
        o2, h2, e2 = R.rnn_encoder(labels, dict(enc.state_dict()))
    assert relerr(o2, out) < 1e-5 and relerr(h2, hid) < 1e-5 and relerr(e2, emb) < 1e-5


def test_predict_chain_vs_reference():
    """dynamic filter -> crop -> res5 (reference's own layer4 as glue) -> mask head, 38x63 map."""
    net = shim.build_reference_net(seed=7)
    g = torch.Generator().manual_seed(8)
    X = torch.relu(torch.randn(1, 1024, 38, 63, generator=g))
    labels, _ = R.synth_labels(g, 1, 10, 1999)
    rois = R.synth_rois(g, 3, 600, 1000)
    with torch.no_grad():
        Y, _, _, _, mask_prob = shim.run_predict(net, X, labels, rois)
        _, hidden, _ = net.rnn_encoder(labels)
        f, w = R.filter_generator(hidden, [getattr(net, "dynamic_fc_%d" % k).weight for k in range(7)],
                                  [getattr(net, "dynamic_fc_%d" % k).bias for k in range(7)],
                                  net.response_fc.weight, net.response_fc.bias)
        r2, Y2 = R.dynamic_filter(X, f, w)
        assert relerr(r2, net._predictions["response"]) < 1e-5
        assert relerr(Y2, Y) < 1e-5
        pool5 = R.crop_pool(Y2, rois)
        fc7 = net.resnet.layer4(pool5)
        s, pr = R.mask_head(fc7, net.mask_up_sampling.weight, net.mask_up_sampling.bias,
                            net.mask_pred_net.weight, net.mask_pred_net.bias)
        assert relerr(s, net._predictions["mask_score"]) < 1e-5
        assert relerr(pr, mask_prob) < 1e-5

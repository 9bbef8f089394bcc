"""Diagnostic: caption-model gradients on the GPU vs the CPU oracle for several (B, L), with and without the
tensor-core Linear for the two big projections."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import restate as R
from lang2seg_b200 import caption_models

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))

for B, L in [(5, 10), (16, 10), (5, 20), (16, 20), (48, 10)]:
    o = dict(vocab_size=1999, input_encoding_size=512, rnn_size=512, att_hid_size=512, fc_feat_size=4096,
             att_feat_size=4096, seq_length=L, num_layers=1, drop_prob_lm=0.5, caption_model="att2in2")
    torch.manual_seed(3)
    model = caption_models.setup(o).cuda().eval()
    g = torch.Generator().manual_seed(11)
    labels, lens = R.synth_labels(g, B, L, o["vocab_size"])
    cap, msk = R.caption_targets(labels, lens, L)
    att0 = torch.relu(torch.randn(B, 14, 14, 4096, generator=g))
    fc = torch.randn(B, 4096, generator=g)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.named_parameters()}
    atto = att0.clone().requires_grad_(True)
    lo = R.caption_loss(fc, atto, cap, msk, params)
    lo.backward()
    for big in (True, False):
        if not big:
            model._big_linear = lambda lin, x: lin(x)      # instance attribute shadows the staticmethod
        model.zero_grad()
        att = att0.cuda().requires_grad_(True)
        loss = model.forward_loss(fc.cuda(), att, cap.cuda(), msk.cuda())
        loss.backward()
        model.__dict__.pop("_big_linear", None)
        worst = sorted(((rel(v.grad, params[k].grad), k) for k, v in model.named_parameters()
                        if not k.endswith("alpha_net.bias")), reverse=True)[:4]
        print("B=%d L=%d tc_linear=%s loss %.2e datt %.2e worst %s" % (B, L, big, rel(loss, lo), rel(att.grad, atto.grad),
              ", ".join("%s %.1e" % (k, e) for e, k in worst)), flush=True)

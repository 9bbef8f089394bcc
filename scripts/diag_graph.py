"""Diagnostic: which pieces of the bench step survive CUDA-graph capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, HotPathStep, make_inputs

wl = WORKLOADS["tiny"]
dev = torch.device("cuda:0")
step = HotPathStep(wl, dev, 1)
d = make_inputs(wl, 1234, dev)
meta = d["_meta"]
net = step.net
for _ in range(2):
    step(d)
torch.cuda.synchronize()


def try_capture(name, fn):
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print("OK   ", name, flush=True)
    except Exception as e:
        print("FAIL ", name, " | ".join(str(e).splitlines()[:3])[:400], flush=True)
        torch.cuda.synchronize()


def f_enc():
    return net.rnn_encoder(d["labels"])

def f_enc_bwd():
    o, h, e = net.rnn_encoder(d["labels"])
    h.sum().backward()

def f_filt():
    _, h, _ = net.rnn_encoder(d["labels"])
    return net._dynamic_filter(d["X"], hidden=h.detach(), expr2img=d["e2i"], resp_target=d["resp_tgt"])

def f_dyn_bwd():
    X = d["X"].detach().requires_grad_(True)
    _, h, _ = net.rnn_encoder(d["labels"])
    y = net._dynamic_filter(X, hidden=h, expr2img=d["e2i"], resp_target=d["resp_tgt"])
    (y.sum() + net._losses["loss_response_per_expr"].sum()).backward()

Ycrop = torch.randn(wl["I"] * wl["EPI"], wl["C"], wl["H"], wl["W"], device=dev, requires_grad=True)


def f_crop():
    p = net._crop_pool_layer(Ycrop, d["rois"], max_pool=False)
    p.sum().backward()

def f_mask():
    fc7 = d["fc7"].detach().requires_grad_(True)
    net._mask_prediction(fc7)
    net._mask_loss(d["mlab"], d["mtgt"]).backward()

def f_cap():
    att = d["att"].detach().requires_grad_(True)
    net._caption_loss(d["fc"], att, d["cap"], d["msk"], steps=meta["steps"]).backward()

def f_opt():
    step.opt.step()

def f_zero():
    step.opt.zero_grad(set_to_none=True)

def f_step():
    step(d)

def f_bilstm_only():
    import lang2seg_b200.functional as F
    B, L, H = 4, 10, 512
    xg = torch.randn(B, L, 8 * H, device=dev, requires_grad=True)
    w1 = torch.randn(4 * H, H, device=dev, requires_grad=True)
    w2 = torch.randn(4 * H, H, device=dev, requires_grad=True)
    lens = torch.tensor([10, 3, 5, 7], device=dev)
    f_bilstm_only.keep = (xg, w1, w2, lens)

def f_bilstm_run():
    import lang2seg_b200.functional as F
    xg, w1, w2, lens = f_bilstm_only.keep
    o, h = F.bilstm(xg, w1 * 0.01, w2 * 0.01, lens)
    h.sum().backward()

def f_embed_bwd():
    e = net.rnn_encoder.embedding(d["labels"])
    e.sum().backward()

def f_mlp_bwd():
    e = net.rnn_encoder.mlp(net.rnn_encoder.embedding(d["labels"]).detach())
    e.sum().backward()

TESTS = dict([("encoder fwd", f_enc), ("encoder fwd+bwd", f_enc_bwd), ("filter+dynfilter fwd", f_filt),
              ("dynfilter fwd+bwd", f_dyn_bwd), ("crop fwd+bwd", f_crop), ("mask head fwd+bwd", f_mask),
              ("caption fwd+bwd", f_cap), ("optimizer step", f_opt), ("zero_grad", f_zero), ("whole step", f_step),
              ("bilstm only", f_bilstm_run), ("embedding bwd", f_embed_bwd), ("mlp bwd", f_mlp_bwd)])
if len(sys.argv) > 1:
    name = sys.argv[1]
    if name == "bilstm only":
        f_bilstm_only()
    try_capture(name, TESTS[name])
else:
    import subprocess
    for name in TESTS:
        r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True)
        lines = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith(("OK", "FAIL"))]
        print(lines[-1] if lines else "?? " + name + " " + r.stderr[-300:], flush=True)

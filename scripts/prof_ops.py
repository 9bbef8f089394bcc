#!/usr/bin/env python
"""Run every hot-path kernel at cfg-2 sizes a few times, standalone -- the target of the ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:'roi_crop|dynfilter|att_step|gemm_bf16x3' \
        -c 40 -o gpurun_out/prof python scripts/prof_ops.py [--reps 2] [--only crop,dyn,mask,att]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lang2seg_b200.functional as F  # noqa: E402
from bench import WORKLOADS, make_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--only", default="dyn,crop,cropmax,mask,att")
    ap.add_argument("--workload", default="cfg2")
    a = ap.parse_args()
    only = set(a.only.split(","))
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    d = make_inputs(wl, 1234, dev)
    I, EPI, C, H, W, Rn, NFG = (wl[k] for k in ("I", "EPI", "C", "H", "W", "R", "NFG"))
    E = I * EPI
    g = torch.Generator(device="cuda").manual_seed(5)
    filt = torch.tanh(torch.randn(E, 7, C, device=dev, generator=g) * 0.1).requires_grad_(True)
    fuse = torch.tanh(torch.randn(E, 7, device=dev, generator=g)).requires_grad_(True)
    for _ in range(a.reps):
        X = d["X"].detach().requires_grad_(True)
        r, Y, rl = F.dynamic_filter(X, filt, fuse, d["e2i"], "sigmoid", d["resp_tgt"])
        if "dyn" in only:
            gy = torch.randn_like(Y) * 1e-3
            torch.autograd.grad([Y, rl], [X, filt, fuse], [gy, torch.ones_like(rl)])
            del gy
        Yd = Y.detach().requires_grad_(True)
        if "crop" in only:
            pool = F.roi_crop(Yd, d["rois"], max_pool=False)
            gp = torch.randn_like(pool) * 1e-4
            torch.autograd.grad(pool, Yd, gp)
            del pool, gp
        if "cropmax" in only:
            pool = F.roi_crop(Yd, d["rois"], max_pool=True)
            gp = torch.randn_like(pool) * 1e-4
            torch.autograd.grad(pool, Yd, gp)
            del pool, gp
        del Y, Yd, r
        if "mask" in only:
            up_w = (torch.randn(2048, 256, 2, 2, device=dev) * 0.01).requires_grad_(True)
            up_b = torch.zeros(256, device=dev, requires_grad=True)
            pw = (torch.randn(81, 256, 1, 1, device=dev) * 0.01).requires_grad_(True)
            pb = torch.zeros(81, device=dev, requires_grad=True)
            fc7 = d["fc7"].detach().requires_grad_(True)
            s, _ = F.mask_head(fc7, up_w, up_b, pw, pb)
            loss = F.mask_bce_loss(s, d["mlab"], d["mtgt"])
            torch.autograd.grad(loss, [fc7, up_w, up_b, pw, pb])
            del s, fc7
        if "maskloss" in only:
            # prediction + mask loss as one node: the fused backward of the training step
            up_w = (torch.randn(2048, 256, 2, 2, device=dev) * 0.01).requires_grad_(True)
            up_b = torch.zeros(256, device=dev, requires_grad=True)
            pw = (torch.randn(81, 256, 1, 1, device=dev) * 0.01).requires_grad_(True)
            pb = torch.zeros(81, device=dev, requires_grad=True)
            fc7 = d["fc7"].detach().requires_grad_(True)
            s, _, loss = F.mask_head_with_loss(fc7, up_w, up_b, pw, pb, d["mlab"], d["mtgt"])
            torch.autograd.grad(loss, [fc7, up_w, up_b, pw, pb])
            del s, fc7
        if "att" in only:
            A, D = 196, 512
            att_h = torch.randn(E, D, device=dev, requires_grad=True)
            feats = torch.randn(E, A, D, device=dev, requires_grad=True)
            p_att = torch.randn(E, A, D, device=dev, requires_grad=True)
            aw = (torch.randn(D, device=dev) * 0.04).requires_grad_(True)
            ab = torch.zeros(1, device=dev, requires_grad=True)
            res, _ = F.attention_step(att_h, feats, p_att, aw, ab)
            torch.autograd.grad(res, [att_h, feats, p_att, aw, ab], torch.randn_like(res))
        if "cap" in only:
            # the caption-side projections that run on the tensor-core Linear (att_embed, ctx2att, logit, i2h)
            for (M, N, K) in [(E * 196, 512, 4096), (E * 196, 512, 512), (E * 11, 2000, 512), (E * 11, 2560, 512)]:
                x = torch.relu(torch.randn(M, K, device=dev)).requires_grad_(True)
                w = (torch.randn(N, K, device=dev) * 0.02).requires_grad_(True)
                b = torch.zeros(N, device=dev, requires_grad=True)
                y = F.linear(x, w, b)
                torch.autograd.grad(y, [x, w, b], torch.randn_like(y))
        if "lin" in only:
            for (M, N, K) in [(48, 3072, 512), (48, 1024, 512), (48, 512, 1024), (48, 512, 3072), (48, 7168, 1024)]:
                x = torch.randn(M, K, device=dev)
                w = torch.randn(N, K, device=dev)
                for _ in range(3):
                    F.linear_small(x, w)
    torch.cuda.synchronize()
    print("prof_ops done")


if __name__ == "__main__":
    main()

"""Diagnostic: F.linear (tcgen05 bf16x3) forward / backward against fp64 for the caption model's shapes, repeated
with fresh and recycled allocator state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lang2seg_b200.functional as F

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))

g = torch.Generator().manual_seed(4)
for rep in range(2):
    for (M, K, N) in [(980, 4096, 512), (3136, 4096, 512), (3136, 512, 512), (9408, 4096, 512), (9408, 512, 512),
                      (528, 512, 2560), (528, 512, 2000), (176, 512, 2560)]:
        x = torch.relu(torch.randn(M, K, generator=g))
        w = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g)
        G = torch.randn(M, N, generator=g)
        junk = torch.full((64, 1024, 1024), 7.0, device="cuda"); del junk      # poison recycled allocator blocks
        xs, ws, bs = (t.cuda().requires_grad_(True) for t in (x, w, b))
        y = F.linear(xs, ws, bs)
        gx, gw, gb = torch.autograd.grad((y * G.cuda()).sum(), [xs, ws, bs])
        xd, wd, bd, Gd = x.double(), w.double(), b.double(), G.double()
        print("rep %d M=%d K=%d N=%d  y %.1e  dx %.1e  dw %.1e  db %.1e" % (
            rep, M, K, N, rel(y, xd @ wd.t() + bd), rel(gx, Gd @ wd), rel(gw, Gd.t() @ xd), rel(gb, Gd.sum(0))), flush=True)

"""Dynamic-filter response layer through the C ABI vs golden vectors (reference output) and the oracle."""
import pytest
import torch

from conftest import relerr
from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 1e-4


def _params(d, prefix="p."):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


@pytest.mark.parametrize("tag", ["a", "b"])
def test_golden(golden, tag):
    import lang2seg_b200.functional as F
    d = golden("dynfilter.npz")
    p = _params(d)
    dyn_w = [p["dynamic_fc_%d.weight" % k].cuda().requires_grad_(True) for k in range(7)]
    dyn_b = [p["dynamic_fc_%d.bias" % k].cuda().requires_grad_(True) for k in range(7)]
    rw = p["response_fc.weight"].cuda().requires_grad_(True)
    rb = p["response_fc.bias"].cuda()
    X = d[tag + ".X"].cuda().requires_grad_(True)
    filt, fuse = R.filter_generator(d[tag + ".hidden"].cuda(), dyn_w, dyn_b, rw, rb)   # plain torch ops on GPU
    r, Y, rl = F.dynamic_filter(X, filt, fuse, resp_target=d[tag + ".tgt"][None].cuda())
    assert relerr(r, d[tag + ".response"]) < TOL
    assert relerr(Y, d[tag + ".Y"]) < TOL
    loss = (Y * d[tag + ".G"].cuda()).sum() + rl.sum()
    gX, g3, g0b, grw = torch.autograd.grad(loss, [X, dyn_w[3], dyn_b[0], rw])
    assert relerr(gX, d[tag + ".dX"]) < TOL
    assert relerr(g3, d[tag + ".d_dyn3_w"]) < TOL
    assert relerr(g0b, d[tag + ".d_dyn0_b"]) < TOL
    assert relerr(grw, d[tag + ".d_resp_w"]) < TOL


@pytest.mark.parametrize("cfg", [dict(I=3, C=96, H=32, W=32, e2i=[0, 0, 1, 1, 1, 2, 2]),
                                 dict(I=2, C=512, H=37, W=62, e2i=[0, 1, 1]),
                                 dict(I=4, C=1024, H=8, W=12, e2i=[0, 2, 2, 3]),      # image 1 has no expression
                                 dict(I=1, C=20, H=9, W=13, e2i=[0]),
                                 # tensor-core kernel (H*W % 4 == 0, C % 32 == 0): the cfg-2 shape, 20 expressions on one
                                 # image (two chunks of the 128 MMA rows), a ragged last pixel tile, a single expression
                                 dict(I=2, C=1024, H=32, W=32, e2i=[0, 0, 0, 1, 1, 1]),
                                 dict(I=2, C=64, H=16, W=16, e2i=[0] * 20 + [1]),
                                 dict(I=2, C=256, H=36, W=35, e2i=[0, 1, 1]),
                                 dict(I=1, C=512, H=32, W=32, e2i=[0]),
                                 # H*W % 4 == 2 (600 x 1000 inputs: 38x63 / 37x62 maps): rows alternate between 16- and 8-byte
                                 # alignment; four expressions on one image = two dfilt chunks; odd C shifts whole images
                                 dict(I=2, C=40, H=38, W=63, e2i=[0, 0, 0, 0, 1]),
                                 dict(I=3, C=33, H=6, W=7, e2i=[0, 1, 1, 2]),
                                 dict(I=2, C=33, H=8, W=8, e2i=[0, 1])])                 # odd C on the float4 path
@pytest.mark.parametrize("gate", ["sigmoid", "linear"])
def test_vs_oracle(cfg, gate):
    import lang2seg_b200.functional as F
    I, C, H, W, e2i = cfg["I"], cfg["C"], cfg["H"], cfg["W"], cfg["e2i"]
    E = len(e2i)
    g = torch.Generator().manual_seed(C + H)
    X = torch.relu(torch.randn(I, C, H, W, generator=g))
    filt = torch.tanh(torch.randn(E, 7, C, generator=g) * 0.5)
    fuse = torch.tanh(torch.randn(E, 7, generator=g))
    if gate == "linear":
        filt = filt / C ** 0.5
    G = torch.randn(E, C, H, W, generator=g)
    Gr = torch.randn(E, 1, H, W, generator=g) * 0.1
    tgt = (torch.rand(E, H, W, generator=g) < 0.3).float()

    def run(fn, dev):
        x, f, w = (t.to(dev).clone().requires_grad_(True) for t in (X, filt, fuse))
        if fn == "oracle":
            r, Y = R.dynamic_filter(x, f, w, e2i, gate)
            rl = R.response_loss(r, tgt)
        else:
            r, Y, rl = F.dynamic_filter(x, f, w, torch.tensor(e2i), gate, tgt.to(dev))
        loss = (Y * G.to(dev)).sum() + (r * Gr.to(dev)).sum() + (rl * torch.arange(1, E + 1, device=dev)).sum()
        return [r, Y, rl] + list(torch.autograd.grad(loss, [x, f, w]))

    ref = run("oracle", "cpu")
    out = run("kernel", "cuda")
    for name, a, b in zip(["r", "Y", "rl", "dX", "dfilt", "dfuse"], out, ref):
        assert relerr(a, b) < TOL, name


def test_partition_bounds_contract():
    """Integer partition boundaries: with X = 1 and one-hot filters r_k is exactly C * M_k."""
    import lang2seg_b200.functional as F
    for H, W, C in [(38, 63, 8), (37, 62, 8), (9, 13, 8), (32, 32, 8), (32, 32, 32), (36, 35, 64)]:   # FFMA and tcgen05 paths
        X = torch.ones(1, C, H, W, device="cuda")
        masks = R.partition_masks(H, W)
        for k in range(7):
            filt = torch.zeros(1, 7, C, device="cuda")
            filt[0, k] = 1.0
            fuse = torch.zeros(1, 7, device="cuda")
            fuse[0, k] = 1.0
            r, _, _ = F.dynamic_filter(X, filt, fuse)
            assert torch.equal(r[0, 0].cpu(), masks[k] * C), (H, W, C, k)


@pytest.mark.parametrize("E,C,Dh", [(48, 1024, 1024), (16, 1024, 1024), (8, 512, 1024), (3, 512, 1024), (70, 64, 32)])
def test_filter_generator_fused_vs_oracle(E, C, Dh):
    """generate_filters: the seven dynamic_fc projections + response_fc as ONE GEMM (fwd, dX, dW) vs 8 torch Linears: on the
    tcgen05 GEMM from 8 expressions on when the stacked weight is large (48 / 16 / 8 rows in mostly empty 128-row tiles),
    the skinny exact-fp32 kernels otherwise."""
    from lang2seg_b200.layers.dynamic_filter import DynamicFilterResponse, generate_filters
    torch.manual_seed(E + C)
    mod = DynamicFilterResponse(Dh, C).cuda()
    g = torch.Generator().manual_seed(1)
    hidden = torch.randn(E, Dh, generator=g)
    Gf, Gw = torch.randn(E, 7, C, generator=g), torch.randn(E, 7, generator=g)
    h = hidden.cuda().requires_grad_(True)
    filt, fuse = generate_filters(h, mod.dynamic_fcs, mod.response_fc)
    ((filt * Gf.cuda()).sum() + (fuse * Gw.cuda()).sum()).backward()
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.named_parameters()}
    ho = hidden.clone().requires_grad_(True)
    fo, wo = R.filter_generator(ho, [p["dynamic_fc_%d.weight" % k] for k in range(7)],
                                [p["dynamic_fc_%d.bias" % k] for k in range(7)], p["response_fc.weight"], p["response_fc.bias"])
    ((fo * Gf).sum() + (wo * Gw).sum()).backward()
    assert relerr(filt, fo) < TOL and relerr(fuse, wo) < TOL
    assert relerr(h.grad, ho.grad) < TOL
    for k, v in mod.named_parameters():
        assert relerr(v.grad, p[k].grad) < TOL, k


def test_expr2img_is_validated_on_the_host():
    """an unsorted / out-of-range host-side expr2img raises instead of silently skipping expressions."""
    import lang2seg_b200.functional as F
    from lang2seg_b200._lib import L2SError
    X = torch.randn(2, 64, 8, 8, device="cuda")
    filt, fuse = torch.randn(3, 7, 64, device="cuda"), torch.randn(3, 7, device="cuda")
    for bad in ([1, 0, 1], [0, 1, 2], [0, 0], [-1, 0, 1]):
        with pytest.raises(L2SError):
            F.dynamic_filter(X, filt, fuse, bad)
    r, Y, _ = F.dynamic_filter(X, filt, fuse, [0, 1, 1])
    assert r.shape == (3, 1, 8, 8) and Y.shape == (3, 64, 8, 8)

"""tcgen05 bf16x3 GEMM and the mask head vs fp64 matmul, golden vectors and the oracle."""
import pytest
import torch

from conftest import relerr
from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 1e-4


def test_gemm_f32_exact():
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(130, 70, generator=g), torch.randn(90, 70, generator=g)
    D = F.gemm_f32(A.cuda(), B.cuda())
    assert relerr(D, A.double() @ B.double().t()) < 1e-6


def test_split_bf16():
    import lang2seg_b200.functional as F
    x = torch.randn(37, 50) * 3
    hi, lo = F.split_bf16(x.cuda(), 64)
    rec = hi.float() + lo.float()
    assert float((rec[:, :50].cpu() - x).abs().max() / x.abs().max()) < 2 ** -15
    assert float(rec[:, 50:].abs().max()) == 0.0


@pytest.fixture
def gemm_shape(request, monkeypatch):
    """L2S_GEMM_SHAPE pins the CTA shape of every tcgen05 GEMM launch (0: 128-row tile / 4 epilogue warps,
    1: 128 rows / 8 epilogue warps, 2: 256 rows / 8 epilogue warps) so that each variant sees every problem shape;
    None leaves the library's own choice."""
    shape = getattr(request, "param", None)
    if shape is None:
        monkeypatch.delenv("L2S_GEMM_SHAPE", raising=False)
    else:
        monkeypatch.setenv("L2S_GEMM_SHAPE", str(shape))
    return shape


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 2048), (300, 96, 40), (1000, 1024, 136), (2600, 520, 264)])
@pytest.mark.parametrize("layout", ["kk", "mm", "mk", "km"])
@pytest.mark.parametrize("gemm_shape", [0, 1, 2], indirect=True)
def test_gemm_bf16x3(M, N, K, layout, gemm_shape):
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    a_mn, b_mn = layout[0] == "m", layout[1] == "m"
    if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
        pytest.skip("row stride not 16-byte aligned for this layout")
    a_hi, a_lo = F.split_bf16((A.t().contiguous() if a_mn else A).cuda())
    b_hi, b_lo = F.split_bf16((B.t().contiguous() if b_mn else B).cuda())
    D = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn)
    assert relerr(D, ref) < 3e-5, "bf16x3 product should be ~1e-5 from fp64"
    if layout == "mm":
        for sk in (3, 0):       # 0: the library chooses the split
            D2 = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn, split_k=sk)
            assert relerr(D2, ref) < 3e-5
    bias = torch.randn(N // 4 if N % 4 == 0 else N, generator=g)
    if N % 4 == 0:
        D3 = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn, epilogue=2, bias=bias.cuda(), bias_div=4)
        assert relerr(D3, torch.relu(ref + bias.double().repeat_interleave(4)[None])) < 3e-5


@pytest.mark.parametrize("gemm_shape", [0, 1, 2], indirect=True)
def test_gemm_bf16x3_many_tiles_per_cta(gemm_shape):
    """More tiles than CTAs: the persistent loop, the smem ring wrap-around and the TMEM accumulator hand-off
    between the MMA issuer and the epilogue warps are all exercised several times per CTA."""
    import lang2seg_b200.functional as F
    M, N, K = 30008, 1000, 200
    g = torch.Generator().manual_seed(7)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ B.double().t()
    a_hi, a_lo = F.split_bf16(A.cuda())
    b_hi, b_lo = F.split_bf16(B.cuda())
    D = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K)
    assert relerr(D, ref) < 3e-5
    D = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, epilogue=4, bias=bias.cuda())
    assert relerr(D, ref + bias.double()[None]) < 3e-5
    D0 = torch.randn(M, N, generator=g)
    D = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, epilogue=1, out=D0.cuda())
    assert relerr(D, ref + D0.double()) < 3e-5


def test_mask_head_golden(golden):
    import lang2seg_b200.functional as F
    d = golden("mask_head.npz")
    x = d["x"].cuda().requires_grad_(True)
    ws = [d[k].cuda().requires_grad_(True) for k in ("up_w", "up_b", "pred_w", "pred_b")]
    score, prob = F.mask_head(x, *ws)
    assert relerr(score, d["score"]) < TOL and relerr(prob, d["prob"]) < TOL
    loss = F.mask_bce_loss(score, d["labels"].cuda(), d["tgt"].cuda())
    assert relerr(loss, d["loss"]) < TOL
    gs = torch.autograd.grad(loss, [x] + ws)
    for gk, k in zip(gs, ("dx", "d_up_w", "d_up_b", "d_pred_w", "d_pred_b")):
        assert relerr(gk, d[k]) < TOL, k


@pytest.mark.parametrize("n,gemm_shape", [(1, None), (8, None), (8, 0), (8, 1), (8, 2)], indirect=["gemm_shape"])
def test_mask_head_full_size_vs_oracle(n, gemm_shape):
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(n)
    x = torch.relu(torch.randn(n, 2048, 7, 7, generator=g))
    up_w, up_b = torch.randn(2048, 256, 2, 2, generator=g) * 0.01, torch.randn(256, generator=g) * 0.01
    pw, pb = torch.randn(81, 256, 1, 1, generator=g) * 0.01, torch.randn(81, generator=g) * 0.01
    labels = torch.randint(1, 81, (n,), generator=g)
    tgt = (torch.rand(n, 14, 14, generator=g) < 0.5).float()
    Gs = torch.randn(n, 81, 14, 14, generator=g) * 1e-3

    def run(dev):
        ts = [t.to(dev).clone().requires_grad_(True) for t in (x, up_w, up_b, pw, pb)]
        if dev == "cpu":
            s, p = R.mask_head(*ts)
            loss = R.mask_loss(s, labels, tgt)
        else:
            s, p = F.mask_head(*ts)
            loss = F.mask_bce_loss(s, labels.to(dev), tgt.to(dev))
        tot = loss + (s * Gs.to(dev)).sum() + (p * Gs.to(dev)).sum()      # dense dscore and dprob paths too
        return [s, p, loss] + list(torch.autograd.grad(tot, ts))

    ref, out = run("cpu"), run("cuda")
    for name, a, b in zip(["score", "prob", "loss", "dx", "d_up_w", "d_up_b", "d_pred_w", "d_pred_b"], out, ref):
        assert relerr(a, b) < TOL, name


@pytest.mark.parametrize("n,Cin,Cmid,ncls", [(8, 2048, 256, 81), (5, 128, 32, 9), (3, 64, 24, 5)])
@pytest.mark.parametrize("extra", [False, True])
def test_mask_head_with_loss_fused_backward(n, Cin, Cmid, ncls, extra):
    """prediction + mask loss as one node: the backward that never builds dscore (l2s_mask_head_bce_bwd) against the
    oracle, and its fallback when the scores have another consumer (`extra`) or Cmid is not supported (24)."""
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(100 + n)
    x = torch.relu(torch.randn(n, Cin, 7, 7, generator=g))
    up_w, up_b = torch.randn(Cin, Cmid, 2, 2, generator=g) * 0.02, torch.randn(Cmid, generator=g) * 0.02
    pw, pb = torch.randn(ncls, Cmid, 1, 1, generator=g) * 0.05, torch.randn(ncls, generator=g) * 0.05
    labels = torch.randint(1, ncls, (n,), generator=g)
    tgt = (torch.rand(n, 14, 14, generator=g) < 0.5).float()
    Gs = torch.randn(n, ncls, 14, 14, generator=g) * 1e-3

    def run(dev):
        ts = [t.to(dev).clone().requires_grad_(True) for t in (x, up_w, up_b, pw, pb)]
        if dev == "cpu":
            s, p = R.mask_head(*ts)
            loss = R.mask_loss(s, labels, tgt)
        else:
            s, p, loss = F.mask_head_with_loss(*ts, labels.to(dev), tgt.to(dev))
        tot = 3.0 * loss
        if extra:
            tot = tot + (s * Gs.to(dev)).sum() + (p * Gs.to(dev)).sum()
        return [s, p, loss] + list(torch.autograd.grad(tot, ts))

    ref, out = run("cpu"), run("cuda")
    for name, a, b in zip(["score", "prob", "loss", "dx", "d_up_w", "d_up_b", "d_pred_w", "d_pred_b"], out, ref):
        assert relerr(a, b) < TOL, name


@pytest.mark.parametrize("gemm_shape", [None, 1, 2], indirect=True)
def test_linear_tc_vs_fp64(gemm_shape):
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(4)
    M, K, N = 1960, 4096, 512
    x = torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    G = torch.randn(M, N, generator=g)
    xs, ws, bs = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = F.linear(xs, ws, bs)
    gx, gw, gb = torch.autograd.grad((y * G.cuda()).sum(), [xs, ws, bs])
    xd, wd, bd, Gd = x.double(), w.double(), b.double(), G.double()
    assert relerr(y, xd @ wd.t() + bd) < 3e-5
    assert relerr(gx, Gd @ wd) < 3e-5
    assert relerr(gw, Gd.t() @ xd) < 3e-5
    assert relerr(gb, Gd.sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------
# bf16 variants (north_star: "bf16 variants within 1e-2"): one tensor pass on the bf16-rounded operands
# ------------------------------------------------------------------------------------------------
BF16_TOL = 1e-2


@pytest.mark.parametrize("layout", ["kk", "mm", "mk", "km"])
def test_gemm_bf16_single_pass(layout):
    import lang2seg_b200.functional as F
    M, N, K = 1000, 1024, 136
    g = torch.Generator().manual_seed(11)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    a_mn, b_mn = layout[0] == "m", layout[1] == "m"
    a_hi, a_lo = F.split_bf16((A.t().contiguous() if a_mn else A).cuda())
    b_hi, b_lo = F.split_bf16((B.t().contiguous() if b_mn else B).cuda())
    with F.precision("bf16"):
        assert F.get_precision() == "bf16"
        D = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn)
        Ds = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn, split_k=0) if layout == "mm" else D
    assert F.get_precision() == "fp32"
    # exactly the product of the bf16-rounded operands (fp32 accumulation), and within 1e-2 of the fp32 product
    exact = a_hi.float().cpu().double() @ b_hi.float().cpu().double().t() if layout == "kk" else None
    if exact is not None:
        assert relerr(D, exact) < 1e-5
    e = relerr(D, ref)
    assert 1e-4 < e < BF16_TOL, "single pass must really be single pass (error %g)" % e
    assert relerr(Ds, ref) < BF16_TOL
    D3 = F.gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, M, N, K, a_mn, b_mn)
    assert relerr(D3, ref) < 3e-5, "the switch must not leak out of the context"


@pytest.mark.parametrize("n", [1, 8])
def test_mask_head_bf16_variant_vs_oracle(n):
    """One tensor pass on bf16-rounded operands.  Forward: within 1e-2 of the fp32 oracle.  Backward: the head is
    piecewise linear, and rounding the operands moves ~0.25 % of the ReLU pre-activations across zero, which changes
    single gradient entries by O(1) of their size whatever the arithmetic behind it -- so the gradients are compared
    (1e-2) with the oracle evaluated on the same bf16-rounded inputs and weights (identical ReLU pattern), i.e. the
    reference computation in bf16 storage, and the forward additionally with the unrounded fp32 oracle."""
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(40 + n)
    x = torch.relu(torch.randn(n, 2048, 7, 7, generator=g))
    up_w, up_b = torch.randn(2048, 256, 2, 2, generator=g) * 0.01, torch.randn(256, generator=g) * 0.01
    pw, pb = torch.randn(81, 256, 1, 1, generator=g) * 0.01, torch.randn(81, generator=g) * 0.01
    labels = torch.randint(1, 81, (n,), generator=g)
    tgt = (torch.rand(n, 14, 14, generator=g) < 0.5).float()
    rb = lambda t: t.bfloat16().float()        # noqa: E731

    def oracle(xx, uw, pww):
        ts = [t.clone().requires_grad_(True) for t in (xx, uw, up_b, pww, pb)]
        s, p = R.mask_head(*ts)
        loss = R.mask_loss(s, labels, tgt)
        return [s, p, loss] + list(torch.autograd.grad(loss, ts))

    ts = [t.cuda().clone().requires_grad_(True) for t in (x, up_w, up_b, pw, pb)]
    with F.precision("bf16"):
        s, p, loss = F.mask_head_with_loss(*ts, labels.cuda(), tgt.cuda())
        out = [s, p, loss] + list(torch.autograd.grad(loss, ts))
    names = ["score", "prob", "loss", "dx", "d_up_w", "d_up_b", "d_pred_w", "d_pred_b"]
    ref32 = oracle(x, up_w, pw)
    for name, a, b in list(zip(names, out, ref32))[:3]:
        assert relerr(a, b) < BF16_TOL, name
    refbf = oracle(rb(x), rb(up_w), rb(pw))
    for name, a, b in zip(names, out, refbf):
        assert relerr(a, b) < BF16_TOL, name + " (oracle on bf16-rounded operands)"


def test_linear_bf16_variant():
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(5)
    M, K, N = 1960, 2048, 512
    x = torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    G = torch.randn(M, N, generator=g)
    xs, ws, bs = (t.cuda().requires_grad_(True) for t in (x, w, b))
    with F.precision("bf16"):
        y = F.linear(xs, ws, bs)
        gx, gw, gb = torch.autograd.grad((y * G.cuda()).sum(), [xs, ws, bs])
    xd, wd, bd, Gd = x.double(), w.double(), b.double(), G.double()
    assert relerr(y, xd @ wd.t() + bd) < BF16_TOL
    assert relerr(gx, Gd @ wd) < BF16_TOL
    assert relerr(gw, Gd.t() @ xd) < BF16_TOL
    assert relerr(gb, Gd.sum(0)) < 1e-5


@pytest.mark.parametrize("R,N,K", [(48, 7176, 1024), (480, 2048, 512), (336, 3072, 512), (7, 128, 64), (50, 70, 130)])
def test_wgrad_f32(R, N, K):
    """dW = dY^T X of the skinny linears / recurrences on the library's FFMA GEMM (register-blocked kernel for 4-aligned
    shapes, the generic one otherwise), bit-identical re-run."""
    import lang2seg_b200.functional as F
    g = torch.Generator().manual_seed(R + N + K)
    dy, x = torch.randn(R, N, generator=g), torch.randn(R, K, generator=g)
    D = F.wgrad_f32(dy.cuda(), x.cuda())
    assert D.shape == (N, K)
    assert relerr(D, dy.double().t() @ x.double()) < 2e-6
    assert torch.equal(D, F.wgrad_f32(dy.cuda(), x.cuda()))

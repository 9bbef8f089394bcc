"""ROI pooling kernels (crop-and-resize and RoI max-pool) through the C ABI vs the oracle / golden vectors."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import relerr
from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 1e-4   # north_star: fp32 outputs and gradients within 1e-4 relative (norm-wise)


def _f():
    import lang2seg_b200.functional as F
    return F


CASES = {"p7": dict(max_pool=False), "p14max": dict(max_pool=True),
         "align7": dict(max_pool=False, align=True), "align14max": dict(max_pool=True, align=True)}


@pytest.mark.parametrize("tag", list(CASES))
def test_crop_golden(golden, tag):
    d = golden("crop.npz")
    kw = CASES[tag]
    imhw = (float(d["im_info"][0, 0]), float(d["im_info"][0, 1])) if kw.get("align") else None
    # C=6 is not a multiple of 4: pad channels (the kernel requires C % 4 == 0) and slice back
    b = torch.zeros(1, 8, 9, 13)
    b[:, :6] = d["bottom"]
    b = b.cuda().requires_grad_(True)
    out = _f().roi_crop(b, d["rois"].cuda(), max_pool=kw["max_pool"], align_im_hw=imhw)
    assert relerr(out[:, :6], d[tag + ".out"]) < TOL
    G = torch.zeros_like(out)
    G[:, :6] = d[tag + ".G"].cuda()
    (gb,) = torch.autograd.grad((out * G).sum(), b)
    assert relerr(gb[:, :6], d[tag + ".dbottom"]) < TOL
    assert float(gb[:, 6:].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(3, 40, 32, 32, 70), (2, 64, 38, 63, 33), (2, 16, 50, 80, 20), (1, 8, 80, 120, 9)])
@pytest.mark.parametrize("max_pool", [False, True, "ranked"])
def test_crop_vs_oracle(shape, max_pool):
    B, C, H, W, N = shape
    ranked = max_pool == "ranked"          # 7x7 crop through the sample-per-lane (ranked) backward kernel
    max_pool = max_pool is True
    g = torch.Generator().manual_seed(B * 1000 + C + H)
    bottom = torch.randn(B, C, H, W, generator=g)
    rois = torch.cat([R.synth_rois(g, N, H * 16, W * 16, b) for b in range(B)])
    rois = rois[torch.randperm(rois.shape[0], generator=g)]          # unsorted batch indices
    bo = bottom.clone().requires_grad_(True)
    ref = R.crop_pool(bo, rois, max_pool=max_pool)
    G = torch.randn(ref.shape, generator=g)
    if max_pool:
        # max-pool routing is discontinuous: where the two best samples of a 2x2 window are closer than the
        # coordinate rounding noise, the winner may legitimately differ -- give those windows no gradient
        with torch.no_grad():
            s14 = R.crop_pool(bottom, rois, max_pool=False, pool=14)
            win = s14.unfold(2, 2, 2).unfold(3, 2, 2).reshape(*ref.shape, 4)
            top2 = win.topk(2, dim=-1).values
            G = G * ((top2[..., 0] - top2[..., 1]) > 1e-3)
    (gref,) = torch.autograd.grad((ref * G).sum(), bo)
    bc = bottom.cuda().requires_grad_(True)
    out = _f().roi_crop(bc, rois.cuda(), max_pool=max_pool, bwd_ranked=ranked)
    assert relerr(out, ref) < TOL
    (gb,) = torch.autograd.grad((out * G.cuda()).sum(), bc)
    assert relerr(gb, gref) < TOL
    # deterministic scatter-add: bit-identical on a second run
    (gb2,) = torch.autograd.grad((_f().roi_crop(bc, rois.cuda(), max_pool=max_pool, bwd_ranked=ranked) * G.cuda()).sum(), bc)
    assert torch.equal(gb, gb2)


@pytest.mark.parametrize("max_pool", [False, True, "ranked"])
def test_crop_degenerate_boxes(max_pool):
    """Zero-area, sub-pixel, out-of-map and whole-map boxes: all 49 samples of a box may fall into one cell
    (collision ranks up to 48 in the backward) or entirely into the zero padding."""
    ranked = max_pool == "ranked"
    max_pool = max_pool is True
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 32, 32, 32
    bottom = torch.randn(B, C, H, W, generator=g)
    boxes = [[0, 100., 100., 100., 100.], [0, 37., 41., 39., 43.], [1, 0., 0., 511., 511.], [1, 600., 600., 900., 900.],
             [0, -50., -50., 20., 20.], [1, 496., 496., 511., 511.], [0, 8., 8., 24., 8.], [1, 16., 300., 16., 420.],
             [0, 255.5, 255.5, 256.5, 256.5], [1, 0., 0., 15., 15.]]
    rois = torch.tensor(boxes * 3)
    bo = bottom.clone().requires_grad_(True)
    ref = R.crop_pool(bo, rois, max_pool=max_pool)
    G = torch.randn(ref.shape, generator=g)
    if max_pool:
        with torch.no_grad():
            s14 = R.crop_pool(bottom, rois, max_pool=False, pool=14)
            win = s14.unfold(2, 2, 2).unfold(3, 2, 2).reshape(*ref.shape, 4)
            top2 = win.topk(2, dim=-1).values
            G = G * ((top2[..., 0] - top2[..., 1]) > 1e-3)
    (gref,) = torch.autograd.grad((ref * G).sum(), bo)
    bc = bottom.cuda().requires_grad_(True)
    out = _f().roi_crop(bc, rois.cuda(), max_pool=max_pool, bwd_ranked=ranked)
    assert relerr(out, ref) < TOL
    (gb,) = torch.autograd.grad((out * G.cuda()).sum(), bc)
    assert relerr(gb, gref) < TOL


def test_crop_index_contract():
    """Corner indices: a one-hot map makes every output reveal which pixel it sampled."""
    g = torch.Generator().manual_seed(3)
    H, W = 32, 32
    rois = R.synth_rois(g, 40, 512, 512)
    px, py = R.crop_sample_coords(rois.numpy(), H, W, 7)
    bottom = torch.zeros(1, 4, H, W)
    bottom[0, 0] = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)       # value = x
    bottom[0, 1] = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)       # value = y
    out = _f().roi_crop(bottom.cuda(), rois.cuda()).cpu().double().numpy()
    inside = (px >= 0) & (px <= W - 1)
    got_x = out[:, 0][:, 0, :]                       # row i = 0: interpolated x coordinate
    ok = inside & (py[:, :1] >= 0) & (py[:, :1] <= H - 1)
    assert np.max(np.abs(got_x - px)[ok]) < 1e-4     # interpolating f(x)=x returns px itself


def test_crop_empty_and_errors():
    F = _f()
    from lang2seg_b200._lib import L2SError
    b = torch.randn(1, 8, 16, 16, device="cuda", requires_grad=True)
    out = F.roi_crop(b, torch.zeros(0, 5, device="cuda"))
    assert out.shape == (0, 8, 7, 7)
    with pytest.raises(L2SError):
        F.roi_crop(torch.randn(1, 6, 16, 16, device="cuda"), torch.zeros(1, 5, device="cuda"))   # C % 4 != 0
    with pytest.raises(AssertionError):
        F.roi_crop(torch.randn(1, 8, 16, 16), torch.zeros(1, 5))                                  # CPU tensors


@pytest.mark.parametrize("shape", [(2, 5, 12, 17, 9), (1, 64, 38, 63, 50), (3, 32, 32, 32, 40)])
def test_roi_maxpool_bit_exact(shape, monkeypatch):
    B, C, H, W, N = shape
    g = torch.Generator().manual_seed(C + N)
    f = torch.randn(B, C, H, W, generator=g)
    rois = torch.cat([R.synth_rois(g, N, H * 16, W * 16, b) for b in range(B)])
    rois = torch.cat([rois, torch.tensor([[0, 5000., 5000., 6000., 6000.], [0, 40., 40., 40., 40.],
                                          [0, 100., 90., 30., 20.]])])
    ref, refarg = R.roi_max_pool(f, rois)
    fc = f.cuda().requires_grad_(True)
    out, arg = _f().roi_max_pool(fc, rois.cuda(), 7, 7, 1.0 / 16, return_argmax=True)
    assert torch.equal(out.cpu(), ref)               # integer bins + fp32 max: bit exact
    assert torch.equal(arg.cpu(), refarg)
    top = torch.randn(ref.shape, generator=g)
    gref = R.roi_max_pool_backward(top, rois, refarg, f.shape)
    (gb,) = torch.autograd.grad((out * top.cuda()).sum(), fc, retain_graph=True)
    assert relerr(gb, gref) < 1e-6                   # default: atomic scatter
    monkeypatch.setenv("L2S_ROIPOOL_BWD_DETERMINISTIC", "1")
    (gd,) = torch.autograd.grad((out * top.cuda()).sum(), fc, retain_graph=True)
    (gd2,) = torch.autograd.grad((out * top.cuda()).sum(), fc)
    assert relerr(gd, gref) < 1e-6
    assert torch.equal(gd, gd2), "owner-computes backward: fixed summation order, bit-identical re-run"


def test_roi_maxpool_vs_reference_cuda_kernel():
    """Bit-exact against the reference's own ROIPoolForward/Backward kernels (roi_pooling_kernel.cu),
    compiled for sm_100a by oracle/Makefile into oracle/_ref (batch 1, as the reference requires)."""
    from oracle import clib
    lib = clib.reference_roi_pool_cuda()
    if lib is None:
        pytest.skip("oracle/_ref/libroi_pooling_ref.so not built")
    g = torch.Generator().manual_seed(17)
    C, H, W, N = 96, 38, 63, 120
    f = torch.randn(1, C, H, W, generator=g).cuda()
    rois = R.synth_rois(g, N, H * 16, W * 16).cuda()
    out_ref = torch.zeros(N, C, 7, 7, device="cuda")
    arg_ref = torch.zeros(N, C, 7, 7, device="cuda", dtype=torch.int32)
    vp = ctypes.c_void_p
    st = vp(torch.cuda.current_stream().cuda_stream)
    lib.ROIPoolForwardLaucher(vp(f.data_ptr()), ctypes.c_float(1 / 16.), N, H, W, C, 7, 7, vp(rois.data_ptr()),
                              vp(out_ref.data_ptr()), vp(arg_ref.data_ptr()), st)
    out, arg = _f().roi_max_pool(f, rois, 7, 7, 1.0 / 16, return_argmax=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref) and torch.equal(arg, arg_ref)
    top = torch.randn(N, C, 7, 7, generator=g).cuda()
    gb_ref = torch.zeros(1, C, H, W, device="cuda")
    lib.ROIPoolBackwardLaucher(vp(top.data_ptr()), ctypes.c_float(1 / 16.), 1, N, H, W, C, 7, 7, vp(rois.data_ptr()),
                               vp(gb_ref.data_ptr()), vp(arg_ref.data_ptr()), st)
    fr = f.clone().requires_grad_(True)
    o2 = _f().roi_max_pool(fr, rois)
    (gb,) = torch.autograd.grad((o2 * top).sum(), fr)
    torch.cuda.synchronize()
    assert relerr(gb, gb_ref) < 1e-6


def _maxpool_tie_mask(bottom, rois, ref_shape):
    """1 where the two best samples of a 2x2 window differ by more than the coordinate rounding noise"""
    with torch.no_grad():
        s14 = R.crop_pool(bottom, rois, max_pool=False, pool=14)
        win = s14.unfold(2, 2, 2).unfold(3, 2, 2).reshape(*ref_shape, 4)
        top2 = win.topk(2, dim=-1).values
        return ((top2[..., 0] - top2[..., 1]) > 1e-3)


# the benchmark's own shapes (BASELINE.json configs 2/3 and 5): one expression = all its ROIs on its own map
FULL = {"cfg2_7x7": dict(B=2, C=1024, H=32, W=32, N=256, max_pool=False),
        "cfg3_14max": dict(B=2, C=1024, H=32, W=32, N=256, max_pool=True),
        "cfg5_vgg_14max": dict(B=2, C=512, H=37, W=62, N=300, max_pool=True),
        "cfg5_vgg_7x7": dict(B=1, C=512, H=37, W=62, N=300, max_pool=False)}


@pytest.mark.parametrize("tag", list(FULL))
def test_crop_full_size_vs_oracle(tag):
    """ROI crop at the real channel / ROI / map sizes of the bench workloads against oracle.restate.crop_pool."""
    k = FULL[tag]
    B, C, H, W, N, max_pool = (k[x] for x in ("B", "C", "H", "W", "N", "max_pool"))
    g = torch.Generator().manual_seed(len(tag) * 7 + C)
    bottom = torch.relu(torch.randn(B, C, H, W, generator=g))
    rois = torch.cat([R.synth_rois(g, N, H * 16, W * 16, b) for b in range(B)])
    bo = bottom.clone().requires_grad_(True)
    ref = R.crop_pool(bo, rois, max_pool=max_pool)
    G = torch.randn(ref.shape, generator=g)
    if max_pool:
        G = G * _maxpool_tie_mask(bottom, rois, ref.shape)
    (gref,) = torch.autograd.grad((ref * G).sum(), bo)
    bc = bottom.cuda().requires_grad_(True)
    out = _f().roi_crop(bc, rois.cuda(), max_pool=max_pool)
    assert relerr(out, ref) < TOL
    (gb,) = torch.autograd.grad((out * G.cuda()).sum(), bc)
    assert relerr(gb, gref) < TOL


@pytest.mark.parametrize("max_pool", [False, True])
def test_crop_bench_scale_properties(max_pool):
    """cfg-2 at full size (48 expressions x 256 ROIs, C = 1024): size-independent properties of the 32-chunk grid --
    expressions are independent (a random pair of them equals the oracle on those two alone), the backward is
    bit-reproducible, and forward/backward are adjoint: <crop(Y), G> == <Y, crop^T(G)>."""
    E, C, H, W, N = 48, 1024, 32, 32, 256
    g = torch.Generator().manual_seed(77)
    Y = torch.relu(torch.randn(E, C, H, W, generator=g)).cuda().requires_grad_(True)
    rois = torch.cat([R.synth_rois(g, N, H * 16, W * 16, e) for e in range(E)]).cuda()
    out = _f().roi_crop(Y, rois, max_pool=max_pool)
    G = torch.randn(out.shape, generator=torch.Generator(device="cuda").manual_seed(3), device="cuda")
    (gY,) = torch.autograd.grad((out * G).sum(), Y, retain_graph=True)
    (gY2,) = torch.autograd.grad((out * G).sum(), Y)
    assert torch.equal(gY, gY2)
    lhs = float((out.double() * G.double()).sum())
    rhs = float((Y.detach().double() * gY.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-3      # piecewise linear in Y (max-pool: same winners)
    for e in (5, 41):
        sl = slice(e * N, (e + 1) * N)
        r1 = rois[sl].cpu().clone()
        r1[:, 0] = 0
        ye = Y[e:e + 1].detach().cpu().requires_grad_(True)
        ref = R.crop_pool(ye, r1, max_pool=max_pool)
        assert relerr(out[sl], ref) < TOL
        Ge = G[sl].cpu()
        if max_pool:
            keep = _maxpool_tie_mask(ye.detach(), r1, ref.shape)
            # compare the gradient only through windows without near-ties: re-run both sides with the masked G
            Ge = Ge * keep
            Gm = torch.zeros_like(G)
            Gm[sl] = Ge.cuda()
            (gsel,) = torch.autograd.grad((_f().roi_crop(Y, rois, max_pool=True) * Gm).sum(), Y)
            got = gsel[e:e + 1]
        else:
            got = gY[e:e + 1]
        (gref,) = torch.autograd.grad((ref * Ge).sum(), ye)
        assert relerr(got, gref) < TOL


@pytest.mark.parametrize("max_pool", [False, True, "ranked"])
def test_crop_bwd_workspace_reuse_is_bit_identical(max_pool):
    """The autograd path hands the forward's workspace (ROI binning + geometry records) to the backward
    (L2S_CROP_WS_PREPARED); a stand-alone l2s_roi_crop_bwd call with a fresh workspace must give the same bits."""
    F = _f()
    from lang2seg_b200 import _lib
    from lang2seg_b200._lib import call, ptr, stream
    ranked = max_pool == "ranked"
    mp = max_pool is True
    B, C, H, W, N = 3, 64, 32, 32, 40
    g = torch.Generator().manual_seed(77)
    bottom = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    rois = torch.cat([R.synth_rois(g, N, H * 16, W * 16, b) for b in range(B)])
    rois = rois[torch.randperm(rois.shape[0], generator=g)].cuda()
    out = F.roi_crop(bottom, rois, max_pool=mp, bwd_ranked=ranked)
    G = torch.randn(out.shape, generator=g).cuda()
    (gb,) = torch.autograd.grad((out * G).sum(), bottom)
    flags = (F.CROP_MAX_POOL if mp else 0) | (F.CROP_BWD_RANKED if ranked else 0)
    arg = None
    nb = _lib.size("l2s_roi_crop_workspace_bytes", B, rois.shape[0], flags)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    if mp:      # the stand-alone call needs the winners: rerun the forward through the C ABI
        arg = torch.empty(out.shape, dtype=torch.uint8, device="cuda")
        out2 = torch.empty_like(out)
        call("l2s_roi_crop_fwd", ptr(bottom.detach()), ptr(rois), ptr(out2), ptr(arg), B, C, H, W, rois.shape[0], 7, flags,
             0.0, 0.0, ptr(ws), nb, stream())
        assert torch.equal(out2, out)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    gb2 = torch.empty_like(gb)
    call("l2s_roi_crop_bwd", ptr(G), ptr(rois), ptr(arg), ptr(gb2), B, C, H, W, rois.shape[0], 7, flags, 0.0, 0.0, ptr(ws), nb,
         stream())
    assert torch.equal(gb, gb2)

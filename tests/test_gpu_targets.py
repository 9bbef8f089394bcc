"""Mask / response targets built on the device vs the oracle (bit exact) and the reference's imresize golden."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def test_resize_masks_golden(golden):
    import lang2seg_b200.functional as F
    d = golden("imresize.npz")
    m = d["mask"].to(torch.uint8)[None].cuda()
    assert torch.equal(F.resize_masks_nearest(m, 9, 13)[0].cpu(), d["r9x13"].float())
    assert torch.equal(F.resize_masks_nearest(m, 32, 32)[0].cpu(), d["r32x32"].float())


@pytest.mark.parametrize("imH,imW,G,n", [(512, 512, 3, 64), (375, 500, 2, 33), (37, 53, 1, 10)])
def test_mask_targets_bit_exact(imH, imW, G, n):
    import lang2seg_b200.functional as F
    g = np.random.RandomState(imH + n)
    masks = (g.rand(G, imH, imW) < 0.45).astype(np.uint8)
    x1 = g.uniform(0, imW - 2, n)
    y1 = g.uniform(0, imH - 2, n)
    x2 = np.minimum(x1 + g.uniform(0, imW * 0.6, n), imW - 1)
    y2 = np.minimum(y1 + g.uniform(0, imH * 0.6, n), imH - 1)
    rois = np.stack([np.zeros(n), x1, y1, x2, y2], 1).astype(np.float32)
    rois[0, 1:] = [0, 0, imW - 1, imH - 1]               # whole image
    rois[1, 1:] = [5.7, 3.2, 5.9, 3.4]                   # a single pixel
    rois[2, 1:] = [imW - 1, imH - 1, imW - 1, imH - 1]   # the last pixel
    assign = g.randint(0, G, n)
    ref = R.mask_targets(masks, rois, assign, 14)
    out = F.mask_targets(torch.from_numpy(masks).cuda(), torch.from_numpy(rois).cuda(), torch.from_numpy(assign).cuda(), 14)
    assert torch.equal(out.cpu(), torch.from_numpy(ref))
    # response target of the same masks at feature resolution
    rt = F.resize_masks_nearest(torch.from_numpy(masks).cuda(), 32, 32).cpu().numpy()
    for k in range(G):
        assert np.array_equal(rt[k], R.nearest_resize_mask(masks[k], 32, 32))


def test_mask_targets_empty_and_errors():
    import lang2seg_b200.functional as F
    m = torch.zeros(1, 8, 8, dtype=torch.uint8, device="cuda")
    out = F.mask_targets(m, torch.zeros(0, 5, device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert out.shape == (0, 14, 14)
    with pytest.raises(RuntimeError):
        F.mask_targets(m, torch.zeros(2, 4, device="cuda"), torch.zeros(2, dtype=torch.int64, device="cuda"))

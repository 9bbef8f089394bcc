#!/usr/bin/env python
"""One forward + backward of every ROI-crop kernel variant at small sizes -- the target of
`compute-sanitizer --tool racecheck|memcheck python scripts/sanitize_crop.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lang2seg_b200.functional as F  # noqa: E402
from lang2seg_b200 import synth  # noqa: E402


def main():
    g = torch.Generator().manual_seed(3)
    for (E, C, H, W, R, maxpool, ranked) in [(2, 64, 32, 32, 24, False, False),    # row-owner 7x7 (roi_crop_bwd_rows_kernel)
                                             (2, 64, 32, 32, 24, True, False),     # generic row-owner, 14x14 + 2x2 max
                                             (1, 32, 37, 62, 20, True, False),     # generic row-owner, 16-channel accumulators
                                             (1, 32, 37, 62, 20, False, False),    # 7x7 on a large map
                                             (2, 64, 32, 32, 24, False, True),     # ranked (sample-per-lane) kernel
                                             (2, 64, 32, 32, 24, True, True)]:
        Y = torch.randn(E, C, H, W, generator=g).cuda().requires_grad_(True)
        rois = torch.cat([synth.synth_rois(g, R, H * 16, W * 16, e) for e in range(E)]).cuda()
        pool = F.roi_crop(Y, rois, max_pool=maxpool, bwd_ranked=ranked)
        (gy,) = torch.autograd.grad(pool, Y, torch.randn_like(pool))
        torch.cuda.synchronize()
        print("ok", E, C, H, W, R, maxpool, ranked, float(pool.abs().sum()), float(gy.abs().sum()))


if __name__ == "__main__":
    main()

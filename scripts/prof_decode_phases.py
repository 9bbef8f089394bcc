#!/usr/bin/env python
"""Phase timeline of the persistent decode kernels (CTA 0, %globaltimer marks through l2s_set_debug_buffer).

    python scripts/prof_decode_phases.py [--B 48] [--L 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lang2seg_b200 import _lib, caption_models  # noqa: E402
from lang2seg_b200 import synth as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=48)
    ap.add_argument("--L", type=int, default=10)
    a = ap.parse_args()
    B, L = a.B, a.L
    o = dict(vocab_size=1999, input_encoding_size=512, rnn_size=512, att_hid_size=512, fc_feat_size=4096,
             att_feat_size=4096, seq_length=L, num_layers=1, drop_prob_lm=0.5, caption_model="att2in2")
    torch.manual_seed(0)
    model = caption_models.setup(o).cuda().eval()
    g = torch.Generator().manual_seed(1)
    labels, lens = R.synth_labels(g, B, L, 1999)
    cap, msk = R.caption_targets(labels, lens, L)
    att = torch.relu(torch.randn(B, 14, 14, 4096, generator=g)).cuda().requires_grad_(True)
    fc = torch.randn(B, 4096, generator=g).cuda()
    buf = torch.zeros(2 * 64 * 8, dtype=torch.int64, device="cuda")
    for it in range(3):
        if it == 2:
            _lib.call("l2s_set_debug_buffer", _lib.ptr(buf), buf.numel() * 8)
        loss = model.forward_loss(fc, att, cap.cuda(), msk.cuda(), steps=int(lens.max()) + 1)
        loss.backward()
        torch.cuda.synchronize()
    _lib.call("l2s_set_debug_buffer", None, 0)
    T = int(lens.max()) + 1
    t = buf.cpu().view(2, 64, 8)
    names = {0: ["start", "A1 done", "A2 done", "bar1 passed", "attention done", "bar2 passed", "C+gates done", "bar3 passed"],
             1: ["start", "S1 done", "bar1 passed", "S2+reduce done", "S4a done", "bar2 passed", "S3 attention done", "bar3 passed"]}
    for k, tag in ((0, "forward"), (1, "backward")):
        print("== decode %s  B=%d T=%d  (us since the step's start mark, CTA 0)" % (tag, B, T))
        steps = range(T) if k == 0 else range(T - 1, -1, -1)
        prev_start = None
        for s in steps:
            row = t[k, s].tolist()
            if row[0] == 0:
                continue
            rel = ["%s %.1f" % (names[k][i], (row[i] - row[0]) / 1e3) for i in range(1, 8) if row[i] >= row[0] and row[i] != 0]
            gap = "" if prev_start is None else "  [step period %.1f us]" % ((row[0] - prev_start) / 1e3)
            prev_start = row[0]
            print("  step %2d: %s%s" % (s, " | ".join(rel), gap))


if __name__ == "__main__":
    main()

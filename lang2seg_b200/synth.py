"""Seeded synthetic inputs of the benchmark workloads.  Plain torch on the host; no kernels,
no oracle: bench.py's B200 arm builds its inputs here so that it never touches `oracle/`."""
import torch


def synth_rois(gen, n, im_h, im_w, batch_index=0):
    """x1,y1 ~ U(0,0.8*dim) ; w,h ~ U(16,0.6*dim) ; clipped to [0,dim-1] ; col0 = batch index."""
    x1 = torch.rand(n, generator=gen) * 0.8 * im_w
    y1 = torch.rand(n, generator=gen) * 0.8 * im_h
    w = 16 + torch.rand(n, generator=gen) * (0.6 * im_w - 16)
    h = 16 + torch.rand(n, generator=gen) * (0.6 * im_h - 16)
    x2 = (x1 + w).clamp(0, im_w - 1)
    y2 = (y1 + h).clamp(0, im_h - 1)
    b = torch.full((n,), float(batch_index))
    return torch.stack([b, x1, y1, x2, y2], 1).float()


def synth_labels(gen, B, L, vocab):
    """Tokens U[1,vocab-1], length U[2,L], zero padded, row 0 has full length (lib/layers/lang_encoder.py:45)."""
    lab = torch.randint(1, vocab, (B, L), generator=gen)
    lens = torch.randint(2, L + 1, (B,), generator=gen)
    lens[0] = L
    lab = lab * (torch.arange(L)[None] < lens[:, None])
    return lab.long(), lens


def caption_targets(labels, lens, L):
    """cap_labels = [0,w1..wn,0..] (B,L+2) ; mask 1 up to and including EOS (lib/loaders/cycle_loader.py:300-308)."""
    B = labels.shape[0]
    cap = torch.zeros(B, L + 2, dtype=torch.long)
    cap[:, 1:L + 1] = labels
    msk = (torch.arange(L + 2)[None] < (lens[:, None] + 2)).float()
    return cap, msk

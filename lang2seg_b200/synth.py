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


def chain_batch(gen, I, EPI, C, H, W, R, NFG, L, vocab, num_classes=81):
    """Host inputs of the res5-chained TRAIN step (HotPathNet.chained_train_step): one referred object per expression
    (box + uint8 mask in image pixels, 16 px per C4 cell), R ROIs per expression of which the first NFG are jittered
    copies of the object's box (foreground); the foreground ROIs of ALL expressions come first in `rois`."""
    E = I * EPI
    im_h, im_w = H * 16, W * 16
    labels, lens = synth_labels(gen, E, L, vocab)
    cap, msk = caption_targets(labels, lens, L)
    x1 = torch.rand(E, generator=gen) * 0.5 * im_w
    y1 = torch.rand(E, generator=gen) * 0.5 * im_h
    bw = 0.2 * im_w + torch.rand(E, generator=gen) * 0.3 * im_w
    bh = 0.2 * im_h + torch.rand(E, generator=gen) * 0.3 * im_h
    cls = torch.randint(1, num_classes, (E,), generator=gen).float()
    gt = torch.stack([x1, y1, (x1 + bw).clamp(max=im_w - 1), (y1 + bh).clamp(max=im_h - 1), cls], 1).floor()
    yy, xx = torch.arange(im_h)[None, :, None], torch.arange(im_w)[None, None, :]
    inside = (xx >= gt[:, 0, None, None]) & (xx <= gt[:, 2, None, None]) & (yy >= gt[:, 1, None, None]) & (yy <= gt[:, 3, None, None])
    masks = (inside & (torch.rand(E, im_h, im_w, generator=gen) < 0.8)).to(torch.uint8)
    fg, bg = [], []
    for e in range(E):
        jit = (torch.rand(NFG, 4, generator=gen) - 0.5) * 0.16 * torch.stack([bw[e], bh[e], bw[e], bh[e]])
        b = gt[e, :4][None] + jit
        b[:, 0::2] = b[:, 0::2].clamp(0, im_w - 1)
        b[:, 1::2] = b[:, 1::2].clamp(0, im_h - 1)
        fg.append(torch.cat([torch.full((NFG, 1), float(e)), b], 1))
        bg.append(synth_rois(gen, R - NFG, im_h, im_w, e))
    rois = torch.cat(fg + bg).float()
    roi_labels = torch.cat([cls.repeat_interleave(NFG), torch.zeros(E * (R - NFG))])
    return {"X": torch.relu(torch.randn(I, C, H, W, generator=gen)), "labels": labels,
            "e2i": torch.arange(I).repeat_interleave(EPI).int(), "rois": rois, "roi_labels": roi_labels,
            "gt_boxes": gt.float(), "gt_masks": masks, "cap": cap, "msk": msk,
            "_meta": {"lens": lens.clone(), "steps": int(lens.max()) + 1, "num_fg": E * NFG}}

"""Drop-in for pyutils/mask-faster-rcnn/lib/nms/pth_nms.py: `pth_nms(dets, thresh)` with the greedy scan on the device.

The reference's CUDA branch (pth_nms.py:26-45) sorts by score, calls `_ext.nms.gpu_nms` -- which builds the bit mask on
the GPU, copies it to the host and scans it there (nms_cuda.c:33-62) -- and indexes `order` with the kept positions.
Here both kernels run on the device (l2s_nms); the only host read is the survivor count needed to size the result
(`pth_nms_padded` avoids even that for CUDA-graph capture)."""
import torch

from .. import functional as L2F


def pth_nms(dets, thresh, max_out=0):
    """dets (N,5) [x1,y1,x2,y2,score] CUDA tensor -> indices of the kept boxes, by descending score."""
    assert dets.is_cuda, "lang2seg_b200 kernels need CUDA tensors (no CPU path)"
    order = dets[:, 4].sort(0, descending=True)[1]
    keep, num = L2F.nms_sorted(dets[order].contiguous(), float(thresh), max_out)
    return order[keep[:int(num)]].contiguous()


def pth_nms_padded(dets, thresh, max_out):
    """Same without a host sync: (indices (max_out,) padded with -1, count (1,) int64 on the device)."""
    order = dets[:, 4].sort(0, descending=True)[1]
    keep, num = L2F.nms_sorted(dets[order].contiguous(), float(thresh), max_out)
    idx = torch.arange(max_out, device=dets.device)
    valid = idx < num
    k = keep[:max_out] if keep.numel() >= max_out else torch.cat([keep, keep.new_zeros(max_out - keep.numel())])
    out = torch.where(valid, order[k.clamp(0, max(dets.shape[0] - 1, 0))], torch.full_like(idx, -1))
    return out, num

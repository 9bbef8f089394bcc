"""Device NMS (mask + greedy scan) vs the oracle restatement of the reference's gpu_nms: the keep list is bit exact."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _boxes(rs, n, dim=600, clustered=True):
    if clustered:       # RPN-like: many near-duplicates around a few centres
        k = max(1, n // 20)
        cx, cy = rs.uniform(0, dim, k), rs.uniform(0, dim, k)
        w, h = rs.uniform(16, dim / 2, k), rs.uniform(16, dim / 2, k)
        idx = rs.randint(0, k, n)
        x1 = cx[idx] - w[idx] / 2 + rs.normal(0, 6, n)
        y1 = cy[idx] - h[idx] / 2 + rs.normal(0, 6, n)
        x2 = x1 + w[idx] * rs.uniform(0.8, 1.2, n)
        y2 = y1 + h[idx] * rs.uniform(0.8, 1.2, n)
    else:
        x1, y1 = rs.uniform(0, dim, n), rs.uniform(0, dim, n)
        x2, y2 = x1 + rs.uniform(0, dim / 3, n), y1 + rs.uniform(0, dim / 3, n)
    s = rs.permutation(n).astype(np.float32) / n              # distinct scores: the sort order is unambiguous
    return np.stack([x1, y1, x2, y2, s], 1).astype(np.float32)


@pytest.mark.parametrize("n,thresh,clustered", [(1, 0.7, True), (63, 0.7, True), (64, 0.5, False), (65, 0.7, True),
                                                 (1000, 0.7, True), (3000, 0.3, False), (6000, 0.7, True)])
def test_nms_vs_oracle(n, thresh, clustered):
    import lang2seg_b200.functional as F
    from lang2seg_b200.nms.pth_nms import pth_nms
    rs = np.random.RandomState(n)
    dets = _boxes(rs, n, clustered=clustered)
    order = np.argsort(-dets[:, 4], kind="stable")
    ref = R.nms_sorted(dets[order], thresh)
    keep, num = F.nms_sorted(torch.from_numpy(dets[order]).cuda(), thresh)
    got = keep[:int(num)].cpu().numpy()
    assert np.array_equal(got, ref)
    # the reference-facing wrapper returns indices into the unsorted detections
    assert np.array_equal(pth_nms(torch.from_numpy(dets).cuda(), thresh).cpu().numpy(), order[ref])
    # survivors never overlap more than thresh with an earlier survivor (idempotence of the keep set)
    again, num2 = F.nms_sorted(torch.from_numpy(dets[order][ref]).cuda(), thresh)
    assert int(num2) == len(ref)


def test_nms_max_out_and_empty():
    import lang2seg_b200.functional as F
    from lang2seg_b200.nms.pth_nms import pth_nms_padded
    rs = np.random.RandomState(5)
    dets = _boxes(rs, 2000, clustered=False)
    order = np.argsort(-dets[:, 4], kind="stable")
    ref = R.nms_sorted(dets[order], 0.7)
    keep, num = F.nms_sorted(torch.from_numpy(dets[order]).cuda(), 0.7, max_out=100)
    assert int(num) == 100 and np.array_equal(keep[:100].cpu().numpy(), ref[:100])
    idx, cnt = pth_nms_padded(torch.from_numpy(dets).cuda(), 0.7, 100)
    assert int(cnt) == 100 and np.array_equal(idx.cpu().numpy(), order[ref[:100]])
    keep, num = F.nms_sorted(torch.zeros(0, 5, device="cuda"), 0.7)
    assert int(num) == 0 and keep.numel() == 0


def test_nms_vs_reference_cuda_kernel():
    """Keep list against the REFERENCE's own bit-mask kernel (nms_kernel.cu, compiled for sm_100a by oracle/Makefile into
    oracle/_ref) followed by the host scan of nms_cuda.c:44-56 restated in numpy."""
    import lang2seg_b200.functional as F
    from oracle import clib
    lib = clib.reference_nms_cuda()
    if lib is None:
        pytest.skip("oracle/_ref/libnms_ref.so not built")
    for n, thresh in ((777, 0.7), (4000, 0.5)):
        rs = np.random.RandomState(n)
        dets = _boxes(rs, n, clustered=True)
        order = np.argsort(-dets[:, 4], kind="stable")
        boxes = torch.from_numpy(dets[order]).cuda().contiguous()
        cb = (n + 63) // 64
        mask = torch.zeros(n, cb, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        lib._nms(n, boxes.data_ptr(), mask.data_ptr(), thresh)          # launches on the default stream
        torch.cuda.synchronize()
        m = mask.cpu().numpy().view(np.uint64)
        remv = np.zeros(cb, np.uint64)
        keep_ref = []
        for i in range(n):                                               # nms_cuda.c:44-56
            nb, ib = divmod(i, 64)
            if not (int(remv[nb]) >> ib) & 1:
                keep_ref.append(i)
                remv[nb:] |= m[i, nb:]
        keep, num = F.nms_sorted(boxes, thresh)
        assert np.array_equal(keep[:int(num)].cpu().numpy(), np.asarray(keep_ref, dtype=np.int64))

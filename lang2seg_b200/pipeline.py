"""Host -> device input pipeline for the hot path.

The reference feeds one image at a time with a blocking `.cuda()` per sentence
(MFR/nets/network_cycle_response.py:644); here a batch's tensors are copied from PINNED host memory on a
dedicated copy stream into one of `depth` device-resident slots while the previous batch is still computing,
so that the PCIe transfer of step i+1 overlaps the kernels of step i (CUDA events order the two streams; no
host synchronisation is added).
"""
from __future__ import annotations

import torch


class HostBatchPipeline:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = [dict() for _ in range(depth)]          # name -> device buffer
        self._ready = [torch.cuda.Event() for _ in range(depth)]   # copy finished
        self._free = [None] * depth                           # compute finished with the slot
        self._head = 0       # next slot to fill
        self._tail = 0       # next slot to hand out
        self._inflight = 0
        self.bytes_per_batch = 0

    def submit(self, host_batch):
        """Enqueue the H2D copies of one batch (dict of pinned CPU tensors).  Returns immediately."""
        assert self._inflight < self.depth, "pipeline full: call get()/release() first"
        slot = self._head
        bufs = self._slots[slot]
        cs = self.copy_stream
        if self._free[slot] is not None:
            cs.wait_event(self._free[slot])        # the step that used this slot has finished reading it
        nbytes = 0
        with torch.cuda.stream(cs):
            for k, v in host_batch.items():
                assert v.device.type == "cpu" and v.is_pinned(), "pipeline inputs must be pinned host tensors (%s)" % k
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                b.copy_(v, non_blocking=True)
                nbytes += v.numel() * v.element_size()
            self._ready[slot].record(cs)
        self.bytes_per_batch = nbytes
        self._head = (slot + 1) % self.depth
        self._inflight += 1
        return slot

    def get(self):
        """Device tensors of the oldest submitted batch; the current stream waits for its copy."""
        assert self._inflight > 0, "nothing submitted"
        slot = self._tail
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        self._cur = slot
        return {k: b.detach() for k, b in self._slots[slot].items()}

    def release(self):
        """Mark the batch returned by the last get() as consumed (its slot may be overwritten)."""
        slot = self._cur
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev
        self._tail = (slot + 1) % self.depth
        self._inflight -= 1

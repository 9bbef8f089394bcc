"""Data-parallel plumbing: one process per GPU, images/expressions sharded across ranks, and one
bucketed gradient all-reduce per parameter group.

The reference has no distributed code at all; expressions are independent in forward and backward,
so the only exchange step on the path is the sum of parameter gradients of the filter generator,
the caption model and the heads.  Works with NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_images(expr2img, n_images, rank, world):
    """Images [i0,i1) of this rank and the expressions that belong to them (expressions follow their image,
    so X is read once per image).  expr2img must be non-decreasing.  Returns (i0, i1, e0, e1)."""
    i0, i1 = shard_range(n_images, rank, world)
    e2i = torch.as_tensor(expr2img)
    e0 = int((e2i < i0).sum())
    e1 = int((e2i < i1).sum())
    return i0, i1, e0, e1


class GradBucket:
    """Flat fp32 bucket over a parameter group: pack grads, all-reduce once, unpack."""

    def __init__(self, params, name=""):
        self.params = [p for p in params if p.requires_grad]
        self.name = name
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce(self, async_op=False, average=False, group=None):
        if not self.params:
            return None
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        self._average = average
        self._group = group
        return work

    def unpack(self):
        if not self.params:
            return
        if self._average:
            self.flat.div_(dist.get_world_size(self._group))
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n


class GradientAllReducer:
    """All-reduce of the filter-generator, caption and head gradient groups, launched asynchronously
    (NCCL runs them on its own stream) and waited on before the optimizer step."""

    def __init__(self, groups, average=False, process_group=None):
        self.buckets = [GradBucket(ps, name) for name, ps in groups.items()]
        self.average = average
        self.pg = process_group
        self._works = []

    @property
    def numel(self):
        return sum(b.numel for b in self.buckets)

    def launch(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.pg) == 1:
            return
        self._works = [(b, b.all_reduce(async_op=True, average=self.average, group=self.pg)) for b in self.buckets]

    def wait(self):
        for b, w in self._works:
            if w is not None:
                w.wait()
            b.unpack()
        self._works = []

    def all_reduce(self):
        self.launch()
        self.wait()


class FlatGradients:
    """Gradients of each parameter group live in ONE pre-allocated flat fp32 buffer and `p.grad` are views into it, so
    the all-reduce runs on the flat buffer directly and the fused optimizer reads the views.  Groups can be reduced
    separately, as soon as their last gradient has been written (bench.py: caption + heads overlap the rest of the
    backward).  Two ways to fill the buffers:

      * `zero()` + backward: autograd accumulates into the views in place -- one `add_` kernel per parameter
        (~100 launches, measured 0.23 ms per cfg-2 step);
      * `optimizer.zero_grad(set_to_none=True)` + backward + `pack(names)`: autograd hands over freshly written gradient
        tensors (no accumulate kernels) and ONE fused multi-tensor copy per group moves them into the flat buffer and
        re-points `p.grad` at the views (what bench.py does)."""

    def __init__(self, groups, process_group=None):
        self.pg = process_group
        self.names = list(groups)
        self.flat = {}
        self.params = {}
        self.views = {}
        for name, ps in groups.items():
            ps = [p for p in ps if p.requires_grad]
            self.params[name] = ps
            if not ps:
                continue
            flat = torch.zeros(sum(p.numel() for p in ps), device=ps[0].device, dtype=torch.float32)
            off = 0
            views = []
            for p in ps:
                n = p.numel()
                views.append(flat[off:off + n].view_as(p))
                p.grad = views[-1]
                off += n
            self.flat[name] = flat
            self.views[name] = views

    @property
    def numel(self):
        return sum(f.numel() for f in self.flat.values())

    def zero(self):
        for f in self.flat.values():
            f.zero_()

    def pack(self, names=None):
        """Move the gradients autograd just produced (separate tensors, after zero_grad(set_to_none=True)) into the flat
        buffers with one fused multi-tensor copy per group and make `p.grad` the views again.  A parameter without a
        gradient keeps `grad = None`; its slice of the buffer stays zero."""
        for n in (names or self.names):
            if n not in self.flat:
                continue
            src, dst, ps = [], [], []
            for p, v in zip(self.params[n], self.views[n]):
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    src.append(p.grad); dst.append(v); ps.append(p)
            if src:
                torch._foreach_copy_(dst, src)
                for p, v in zip(ps, dst):
                    p.grad = v

    def all_reduce_async(self, names=None):
        """Launch the (sum) all-reduce of the named groups; returns the work handles (empty without a process group)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.pg) == 1:
            return []
        return [dist.all_reduce(self.flat[n], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                for n in (names or self.names) if n in self.flat]

    @staticmethod
    def wait(works):
        for w in works:
            w.wait()

    def all_reduce(self, names=None):
        self.wait(self.all_reduce_async(names))

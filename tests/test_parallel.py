"""Host-side multi-GPU logic on CPU: sharding and the gloo world_size-2 gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lang2seg_b200.parallel import FlatGradients, GradientAllReducer, shard_images, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 16, 48, 129):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_images_keeps_expressions_with_their_image():
    e2i = [0, 0, 0, 1, 2, 2, 3, 3, 3, 3]
    got = [shard_images(e2i, 4, r, 2) for r in range(2)]
    assert got == [(0, 2, 0, 4), (2, 4, 4, 10)]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin_a, lin_b = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    full = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    lo, hi = shard_range(8, rank, world)
    (lin_b(lin_a(full[lo:hi])) ** 2).sum().backward()
    red = GradientAllReducer({"a": list(lin_a.parameters()), "b": list(lin_b.parameters())})
    red.all_reduce()
    grads = [p.grad.clone() for p in list(lin_a.parameters()) + list(lin_b.parameters())]
    # the same step through flat gradient buffers (p.grad are views; groups reduced separately, asynchronously)
    flat = FlatGradients({"a": list(lin_a.parameters()), "b": list(lin_b.parameters())})
    for _ in range(2):                      # twice: accumulation in place + zero() between steps
        flat.zero()
        (lin_b(lin_a(full[lo:hi])) ** 2).sum().backward()
        wa = flat.all_reduce_async(["b"])
        wb = flat.all_reduce_async(["a"])
        flat.wait(wa + wb)
    flat_grads = [p.grad.clone() for p in list(lin_a.parameters()) + list(lin_b.parameters())]
    assert all(p.grad.data_ptr() >= flat.flat[n].data_ptr() for n, ps in flat.params.items() for p in ps)
    # the pack() protocol: fresh gradients (set_to_none) -> one fused copy per group -> views again
    for _ in range(2):
        for p in list(lin_a.parameters()) + list(lin_b.parameters()):
            p.grad = None
        (lin_b(lin_a(full[lo:hi])) ** 2).sum().backward()
        flat.pack(["b"])
        wb = flat.all_reduce_async(["b"])
        flat.pack(["a"])
        flat.wait(wb + flat.all_reduce_async(["a"]))
    assert all(p.grad.data_ptr() == v.data_ptr() for n in flat.names for p, v in zip(flat.params[n], flat.views[n]))
    packed_grads = [p.grad.clone() for p in list(lin_a.parameters()) + list(lin_b.parameters())]
    assert all(torch.equal(a, b) for a, b in zip(flat_grads, packed_grads))
    # the cut protocol of the N > 1 step (bench.py: Network._dynamic_filter(cut_filters=True) + bwd_rest()): the graph is
    # cut at an intermediate tensor, the downstream group is packed and on the wire while the upstream backward runs
    for _ in range(2):
        for p in list(lin_a.parameters()) + list(lin_b.parameters()):
            p.grad = None
        hidden = lin_a(full[lo:hi])
        leaf = hidden.detach().requires_grad_(True)
        (lin_b(leaf) ** 2).sum().backward()                  # stops at the leaf; lin_b's gradients are complete
        flat.pack(["b"])
        wb = flat.all_reduce_async(["b"])
        torch.autograd.backward([hidden], [leaf.grad])        # the rest of the backward
        flat.pack(["a"])
        flat.wait(wb + flat.all_reduce_async(["a"]))
    cut_grads = [p.grad.clone() for p in list(lin_a.parameters()) + list(lin_b.parameters())]
    assert all(torch.allclose(a, b, rtol=1e-6, atol=1e-7) for a, b in zip(flat_grads, cut_grads))
    if rank == 0:
        torch.save([grads, flat_grads], out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, got_flat = torch.load(out)
    torch.manual_seed(0)
    lin_a, lin_b = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    full = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    (lin_b(lin_a(full)) ** 2).sum().backward()
    ref = [p.grad for p in list(lin_a.parameters()) + list(lin_b.parameters())]
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    for a, b in zip(got_flat, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)

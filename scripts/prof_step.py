#!/usr/bin/env python
"""Per-kernel device time of the bench step measured in situ (warm caches, real overlap) with torch.profiler/CUPTI.

    python scripts/prof_step.py [--workload cfg2] [--steps 3] > gpurun_out/step_kernels.txt
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import WORKLOADS, HotPathStep, make_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--trace", default="", help="comma-separated kernel-name substrings (or 'all'): per-launch timeline of the last step")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    step = HotPathStep(wl, dev, 1)
    d = make_inputs(wl, 1234, dev)
    for _ in range(3):
        step(d)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        step(d)
    t1.record()
    torch.cuda.synchronize()
    print("# un-profiled: %.3f ms/step" % (t0.elapsed_time(t1) / a.steps))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        for _ in range(a.steps):
            step(d)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name[:110]][0] += 1
            agg[ev.name[:110]][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print("# %d steps, %.1f us of kernel time per step (sum over kernels)" % (a.steps, tot / a.steps))
    print("%-110s %8s %12s %7s" % ("kernel", "n/step", "us/step", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        print("%-110s %8.1f %12.1f %6.2f%%" % (k, v[0] / a.steps, v[1] / a.steps, 100 * v[1] / tot))


    # launch-by-launch timeline of the LAST step for the kernels matching --trace (name substring): start offset, duration
    if a.trace:
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda ev: ev.time_range.start)
        per = len(evs) // a.steps
        last = evs[-per:]
        t0 = last[0].time_range.start
        print("\n# timeline of the last step (%d device events), kernels matching %r" % (per, a.trace))
        prev_end = t0
        for ev in last:
            dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            gap = ev.time_range.start - prev_end
            prev_end = max(prev_end, ev.time_range.end)
            if a.trace == "all" or any(p in ev.name for p in a.trace.split(",")):
                print("%10.1f us  +%7.1f us  gap %6.1f  %s" % (ev.time_range.start - t0, dur, gap, ev.name[:120]))

    # framework-side operators by (inclusive) device time, with their input shapes
    print("\n# aten operators by device time (per step, inclusive of children)")
    rows = []
    for e in prof.key_averages(group_by_input_shape=True):
        dt = 0.0
        for attr in ("device_time_total", "cuda_time_total", "self_device_time_total", "self_cuda_time_total"):
            v = getattr(e, attr, None)
            if v:
                dt = float(v)
                break
        if dt > 0 and e.key.startswith(("aten::", "Optimizer", "autograd::")):
            rows.append((dt / a.steps, e.count / a.steps, e.key, str(e.input_shapes)[:120]))
    for dt, n, key, shp in sorted(rows, reverse=True)[:60]:
        print("%9.1f us %6.1f x  %-40s %s" % (dt, n, key[:40], shp))


if __name__ == "__main__":
    main()

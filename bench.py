"""bench.py -- expressions/sec of the lang2seg hot path (fwd+bwd) on B200, next to its CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg1|cfg2|cfg3|cfg4|cfg5|tiny]

A "step" is one forward+backward pass of the chained hot path over one batch of synthetic input
(SURVEY.md section 8d): lang encoder -> filter generator -> dynamic filter (+response BCE) ->
ROI crop (consumes the gated map) -> mask head (+mask BCE, synthetic res5 features) -> att2in2
(+LM loss, synthetic fc/att features), gradient all-reduce of the parameter groups (N>1)
and the SGD update.  res5 is cuDNN glue outside the graded step.

The step is captured in CUDA graphs; its three data-independent branches (encoder -> dynamic filter -> crop | mask
head | att2in2) are issued on separate streams during capture, so they overlap forward and backward (HotPathStep).

N>1 is launched by torchrun (one rank per GPU); every rank processes its own shard of images /
expressions (weak scaling) and only parameter gradients cross NVLink.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own modules (baseline/_ref, under the
compatibility shim; else the oracle's torch-CPU port) on the host cores instead (rank 0 only).

Diagnostics (environment): L2S_BENCH_STREAMS=0|1|2 (branch streams, default 2), L2S_BENCH_FORCE_SPLIT=1 (the N > 1
graph structure on one GPU), L2S_BENCH_NOCOMM=1 (N > 1 without the collectives).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "expressions/sec fwd+bwd (dynfilter+ROIAlign+mask head+att2in2)"
UNIT = "expressions/s"

WORKLOADS = {
    # BASELINE.json configs[0] / SURVEY 8d config 1 -- the reference's own CPU-runnable case, forward only
    "cfg1": dict(name="cfg1: 1 image x 1 expression, C4 1024x32x32, 100 ROIs (mask head on all 100), L=10 (T=11), "
                      "forward only",
                 I=1, EPI=1, C=1024, H=32, W=32, R=100, NFG=100, L=10, V=1999,
                 parts=("resp", "crop7", "mask", "caption"), fwd_only=True),
    # BASELINE.json configs[1] / SURVEY 8d config 2 -- the configuration the metric is quoted on (fits one GPU)
    "cfg2": dict(name="cfg2: 16 images x 3 expressions, C4 1024x32x32, 256 ROIs + 64 fg per expression, "
                      "L=10 (T=11), V=1999, 7x7 crop, response+mask+caption losses",
                 I=16, EPI=3, C=1024, H=32, W=32, R=256, NFG=64, L=10, V=1999,
                 parts=("resp", "crop7", "mask", "caption")),
    # configs[2]: 64 images over 8 GPUs = 8 images x 3 expressions per GPU; the 14x14-sample + 2x2-max crop
    # (network_cycle_response.py:140-144) AND the default 7x7 crop both run on the gated map; mask head; no caption
    "cfg3": dict(name="cfg3 (per-GPU shard of 64 images / 8): 8 images x 3 expressions, C4 1024x32x32, 256 ROIs, "
                      "crop 14x14+2x2max and 7x7, response BCE, mask head 14x14 on 64 fg",
                 I=8, EPI=3, C=1024, H=32, W=32, R=256, NFG=64, L=10, V=1999,
                 parts=("resp", "crop_max", "crop7", "mask")),
    # configs[3]: 128 image/expression pairs over 8 GPUs = 16 per GPU; dynamic filter + att2in2 only (L=20, T=21)
    "cfg4": dict(name="cfg4 (per-GPU shard of 128 / 8): 16 images x 1 expression, C4 1024x32x32, spatial filters + "
                      "att2in2 caption loss, L=20 (T=21), V=1999; no crop, no mask head",
                 I=16, EPI=1, C=1024, H=32, W=32, R=0, NFG=0, L=20, V=1999, parts=("caption", "dY")),
    # configs[4]: VGG16 variant -- conv5_3 512x37x62 (600x1000 image), 300 ROIs, crop with max_pool=True
    # (network_vgg.py:137-141), response loss, no mask head, no caption; --batch sweeps I = E
    "cfg5": dict(name="cfg5 (VGG16): I=E images x 1 expression, conv5_3 512x37x62, 300 ROIs, crop 14x14+2x2max, "
                      "response BCE; no mask head",
                 I=32, EPI=1, C=512, H=37, W=62, R=300, NFG=0, L=10, V=1999, parts=("resp", "crop_max")),
    "tiny": dict(name="tiny smoke workload", I=2, EPI=2, C=1024, H=32, W=32, R=16, NFG=4, L=10, V=1999,
                 parts=("resp", "crop7", "mask", "caption")),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the hot kernels from this round's
    `ncu --set full` capture at cfg-2 sizes (profiles/r01_traffic.json); {} when the file is missing."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        k = json.load(open(p))["kernels"]
        return {name: v["dram_read_bytes"] + v["dram_write_bytes"] for name, v in k.items()}
    except Exception:
        return {}


# bench component -> kernel name in profiles/r01_traffic.json
TRAFFIC_KEY = {"dynfilter_fwd": "dynfilter_fwd", "roi_crop_fwd": "roi_crop_fwd", "roi_crop_bwd": "roi_crop_bwd_rows",
               "gemm_bf16x3_kernel<EpiUp> (mask head GEMM1, in-step epilogue)": "EpiUp", "att_step_fwd": "att_step_fwd"}


def make_inputs(wl, seed, device, pinned=False):
    """Seeded synthetic inputs of SURVEY 8d on the host (only those the workload's parts consume); moved to `device`
    unless pinned host copies are wanted."""
    from lang2seg_b200 import synth as R      # seeded generators (plain torch; the B200 arm never imports oracle/)
    g = torch.Generator().manual_seed(seed)
    I, EPI, C, H, W, Rn, NFG, L, V = (wl[k] for k in ("I", "EPI", "C", "H", "W", "R", "NFG", "L", "V"))
    parts = wl["parts"]
    E = I * EPI
    d = {}
    d["X"] = torch.relu(torch.randn(I, C, H, W, generator=g))
    labels, lens = R.synth_labels(g, E, L, V)
    d["labels"] = labels
    host_meta = {"lens": lens.clone(), "steps": int(lens.max()) + 1}    # host-side facts about the batch (never copied)
    d["e2i"] = torch.arange(I).repeat_interleave(EPI).int()
    if "resp" in parts:
        d["resp_tgt"] = (torch.rand(E, H, W, generator=g) < 0.3).float()
    if "crop7" in parts or "crop_max" in parts:
        d["rois"] = torch.cat([R.synth_rois(g, Rn, H * 16, W * 16, e) for e in range(E)])
    if "mask" in parts:
        d["fc7"] = torch.relu(torch.randn(E * NFG, 2048, 7, 7, generator=g))
        d["mlab"] = torch.randint(1, 81, (E * NFG,), generator=g)
        d["mtgt"] = (torch.rand(E * NFG, 14, 14, generator=g) < 0.5).float()
    if "caption" in parts:
        d["cap"], d["msk"] = R.caption_targets(labels, lens, L)
        d["fc"] = torch.randn(E, 4096, generator=g)
        d["att"] = torch.relu(torch.randn(E, 14, 14, 4096, generator=g))
    if pinned:
        out = {k: v.pin_memory() for k, v in d.items()}
    else:
        out = {k: v.to(device) for k, v in d.items()}
    out["_meta"] = host_meta
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------
class HotPathStep:
    """The chained hot-path step of one workload: which of {response BCE, 7x7 crop, 14x14+max crop, mask head,
    att2in2} run is the workload's `parts`."""

    def __init__(self, wl, device, world):
        from lang2seg_b200.nets.network import HotPathNet
        from lang2seg_b200.parallel import FlatGradients
        torch.manual_seed(1234)
        self.wl, self.device = wl, device
        self.parts = wl["parts"]
        self.fwd_only = bool(wl.get("fwd_only"))
        self.net = HotPathNet(dict(seq_length=wl["L"], vocab_size=wl["V"], C4_feat_dim=wl["C"])).to(device).eval()   # eval: dropout off (D8)
        self.params = [p for p in self.net.parameters() if p.requires_grad]
        self.opt = torch.optim.SGD(self.params, lr=1e-5, momentum=0.9, fused=True)
        # N > 1: gradients are packed into flat per-group buffers by one fused copy per group (p.grad become views)
        self.flat = FlatGradients(self.net.gradient_groups()) if (world > 1 and not self.fwd_only) else None
        self._pending = None
        E = wl["I"] * wl["EPI"]
        g = torch.Generator().manual_seed(99)
        # upstream gradients of pool5 (stand in for res5's backward) and, where nothing else consumes the gated map,
        # of the gated map itself (stands in for layer4 <- caption features): device resident, not inputs
        self.g_pool = self.g_pool_max = self.g_Y = None
        if not self.fwd_only:
            if "crop7" in self.parts:
                self.g_pool = (torch.randn(E * wl["R"], wl["C"], 7, 7, generator=g) * 1e-4).to(device)
            if "crop_max" in self.parts:
                self.g_pool_max = (torch.randn(E * wl["R"], wl["C"], 7, 7, generator=g) * 1e-4).to(device)
            if "dY" in self.parts:
                self.g_Y = (torch.randn(E, wl["C"], wl["H"], wl["W"], generator=g) * 1e-4).to(device)
        self.one = torch.ones((), device=device)
        # Independent branches of the step on their own streams (forward and -- through autograd's stream semantics --
        # backward), so that the captured CUDA graph has parallel branches: the att2in2 branch is latency bound
        # (persistent recurrences on 128 CTAs with 64 KB of shared memory each) and shares the SMs with bandwidth-bound
        # kernels of the other branches; the mask head overlaps the encoder / filter-generator chain and the crops' tails.
        # Measured on one box (profiles/r02_ab.md): cfg2 9.30 -> 9.05 ms/step, cfg3 8.28 -> 7.86, cfg4 2.29 -> 2.01.
        # Only under graph capture: eager launches (the e2e leg) pay more for the stream switches than they gain.
        # L2S_BENCH_STREAMS=0 keeps everything on one stream.
        nstreams = int(os.environ.get("L2S_BENCH_STREAMS", "2"))
        # (stream priorities for the branches were measured and changed nothing: profiles/r02_ab.md)
        self.side = torch.cuda.Stream(device) if (nstreams >= 1 and "caption" in self.parts and not self.fwd_only) else None
        self.side2 = torch.cuda.Stream(device) if (nstreams >= 2 and "mask" in self.parts and not self.fwd_only) else None

    def forward_only(self, d, meta):
        """cfg-1: the inference-side forward (network_cycle_response.py:576-596 without res5) -> a scalar checksum"""
        net = self.net
        with torch.no_grad():
            gated = net._dynamic_filter(d["X"], d["labels"], expr2img=d["e2i"], resp_target=d.get("resp_tgt"),
                                        lengths=meta.get("lens"))
            acc = net._predictions["response"].sum()
            if "crop7" in self.parts:
                acc = acc + net._crop_pool_layer(gated, d["rois"], max_pool=False).sum()
            if "crop_max" in self.parts:
                acc = acc + net._crop_pool_layer(gated, d["rois"], max_pool=True).sum()
            if "mask" in self.parts:
                acc = acc + net._mask_prediction(d["fc7"]).sum()
            if "caption" in self.parts:
                acc = acc + net.caption_model(d["fc"], d["att"], d["cap"], steps=meta.get("steps")).sum()
        net._predictions.clear()
        net._losses.clear()
        return acc

    def fwd_bwd(self, d, meta=None, split=False):
        """forward + backward of the chained hot path; leaves the gradients in p.grad.
        split=True (N > 1): the backward runs down to the generated filters here (the caption / head gradient groups are
        then complete and are all-reduced while `bwd_rest()` runs the remaining backward: filter generator -> language
        encoder)."""
        net, parts = self.net, self.parts
        meta = meta if meta is not None else d.get("_meta", {})
        if self.fwd_only:
            return self.forward_only(d, meta)
        self.opt.zero_grad(set_to_none=True)      # N > 1: the fresh gradients are packed into the flat buffers below
        X = d["X"].requires_grad_(True)
        cap_loss = att = None
        branch = torch.cuda.is_current_stream_capturing()
        if self.side is not None and branch:
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                att = d["att"].requires_grad_(True)
                cap_loss = net._cap_loss_weight * net._caption_loss(d["fc"], att, d["cap"], d["msk"], steps=meta.get("steps"))
        mask_loss = fc7 = None
        if self.side2 is not None and branch:
            cur = torch.cuda.current_stream()
            self.side2.wait_stream(cur)
            with torch.cuda.stream(self.side2):
                fc7 = d["fc7"].requires_grad_(True)
                net._mask_prediction(fc7, d["mlab"], d["mtgt"])
                mask_loss = net._mask_loss(d["mlab"], d["mtgt"])
        gated = net._dynamic_filter(X, d["labels"], expr2img=d["e2i"], resp_target=d.get("resp_tgt"),
                                    lengths=meta.get("lens"), cut_filters=split)
        loss_a, loss_b = 0, 0
        outs, grads = [], []
        if "resp" in parts:
            loss_b = loss_b + net._losses["loss_response_per_expr"].sum()
        if "crop_max" in parts:
            outs.append(net._crop_pool_layer(gated, d["rois"], max_pool=True)); grads.append(self.g_pool_max)
        if "crop7" in parts:
            outs.append(net._crop_pool_layer(gated, d["rois"], max_pool=False)); grads.append(self.g_pool)
        if "dY" in parts:
            outs.append(gated); grads.append(self.g_Y)
        if mask_loss is not None:
            torch.cuda.current_stream().wait_stream(self.side2)
            mask_loss.record_stream(torch.cuda.current_stream())
            loss_a = loss_a + mask_loss
        elif "mask" in parts:
            fc7 = d["fc7"].requires_grad_(True)
            net._mask_prediction(fc7, d["mlab"], d["mtgt"])   # prediction + mask loss as one node (fused backward)
            loss_a = loss_a + net._mask_loss(d["mlab"], d["mtgt"])
        if "caption" in parts and cap_loss is None:
            att = d["att"].requires_grad_(True)
            loss_a = loss_a + net._cap_loss_weight * net._caption_loss(d["fc"], att, d["cap"], d["msk"],
                                                                       steps=meta.get("steps"))
        elif cap_loss is not None:
            torch.cuda.current_stream().wait_stream(self.side)
            cap_loss.record_stream(torch.cuda.current_stream())
            loss_a = loss_a + cap_loss
        roots_a = ([loss_a], [self.one]) if torch.is_tensor(loss_a) else ([], [])
        roots_b = (([loss_b] if torch.is_tensor(loss_b) else []) + outs, ([self.one] if torch.is_tensor(loss_b) else []) + grads)
        loss = (loss_a + loss_b).detach()
        if split:
            # everything except the language encoder: the backward stops at the expression embedding (its gradient is
            # kept), so that the branches still overlap inside this graph and the all-reduce of the caption / head /
            # filter-generator groups overlaps the rest
            torch.autograd.backward(roots_a[0] + roots_b[0], roots_a[1] + roots_b[1])
            stops = net._predictions.get("graph_cut")              # ((hidden,), (its detached leaf,)) or None
            self.flat.pack(self.EARLY_GROUPS)
            self._pending = (stops, X, fc7, att)
        else:
            torch.autograd.backward(roots_a[0] + roots_b[0], roots_a[1] + roots_b[1])
            if self.flat is not None:
                self.flat.pack()
            self._finish(X, fc7, att)
        return loss

    # ---- N > 1, concurrent-graph structure: every branch is its own CUDA graph (forward + backward + gradient pack),
    # replayed on its own stream; the all-reduce of a branch's gradient group is launched behind that branch alone
    def _zero(self, names):
        for n in names:
            for p in self.flat.params.get(n, []):
                p.grad = None

    def branch_caption(self, d, meta=None):
        meta = meta if meta is not None else d.get("_meta", {})
        net = self.net
        self._zero(["caption"])
        att = d["att"].requires_grad_(True)
        loss = net._cap_loss_weight * net._caption_loss(d["fc"], att, d["cap"], d["msk"], steps=meta.get("steps"))
        torch.autograd.backward([loss], [self.one])
        self.flat.pack(["caption"])
        att.grad = None
        return loss.detach()

    def branch_mask(self, d):
        net = self.net
        self._zero(["heads"])
        fc7 = d["fc7"].requires_grad_(True)
        net._mask_prediction(fc7, d["mlab"], d["mtgt"])
        loss = net._mask_loss(d["mlab"], d["mtgt"])
        torch.autograd.backward([loss], [self.one])
        self.flat.pack(["heads"])
        fc7.grad = None
        for k in ("mask_score", "mask_prob"):
            net._predictions.pop(k, None)
        net._losses.pop("mask_loss", None)
        return loss.detach()

    def branch_main(self, d, meta=None):
        """language encoder -> filter generator -> dynamic filter -> ROI crops, backward down to the expression
        embedding (bwd_rest() finishes the encoder)"""
        meta = meta if meta is not None else d.get("_meta", {})
        net, parts = self.net, self.parts
        self._zero(["filter_generator", "encoder"])
        X = d["X"].requires_grad_(True)
        gated = net._dynamic_filter(X, d["labels"], expr2img=d["e2i"], resp_target=d.get("resp_tgt"),
                                    lengths=meta.get("lens"), cut_filters=True)
        roots, grads = [], []
        loss = torch.zeros((), device=self.device)
        if "resp" in parts:
            loss = net._losses["loss_response_per_expr"].sum()
            roots.append(loss); grads.append(self.one)
        if "crop_max" in parts:
            roots.append(net._crop_pool_layer(gated, d["rois"], max_pool=True)); grads.append(self.g_pool_max)
        if "crop7" in parts:
            roots.append(net._crop_pool_layer(gated, d["rois"], max_pool=False)); grads.append(self.g_pool)
        if "dY" in parts:
            roots.append(gated); grads.append(self.g_Y)
        torch.autograd.backward(roots, grads)
        self.flat.pack(["filter_generator"])
        self._pending = (net._predictions.get("graph_cut"), X, None, None)
        return loss.detach()

    def bwd_rest(self):
        """second half of a split backward (see fwd_bwd)"""
        stops, X, fc7, att = self._pending
        self._pending = None
        if stops:
            torch.autograd.backward(list(stops[0]), [t.grad for t in stops[1]])
        self.flat.pack([n for n in self.flat.names if n not in self.EARLY_GROUPS])
        self._finish(X, fc7, att)

    def _finish(self, X, fc7, att):
        X.grad = None
        if fc7 is not None:
            fc7.grad = None
        if att is not None:
            att.grad = None
        # drop the references into this step's autograd graph (the reference keeps them in _predictions/_losses for its
        # tensorboard summaries): a graph kept alive across steps pins AccumulateGrad nodes to the stream they were
        # created on, which breaks CUDA-graph capture on another stream
        self.net._predictions.clear()
        self.net._losses.clear()

    # gradient groups that are complete after the first half of a split backward
    EARLY_GROUPS = ("caption", "heads", "filter_generator")

    def update(self):
        """gradient all-reduce of the parameter groups (N > 1) and the SGD update"""
        if self.fwd_only:
            return
        if self.flat is not None:
            self.flat.all_reduce()
        self.opt.step()

    def __call__(self, d, meta=None):
        loss = self.fwd_bwd(d, meta)
        self.update()
        return loss


class ChainedStep:
    """The reference-faithful chain WITH res5 in the middle (SURVEY 8d "second number", row a8): dynamic filter ->
    7x7 crop -> res5 (cuDNN glue, lang2seg_b200/nets/res5_glue.py) -> box head + mask head ; ungated / gated map ->
    res5 -> caption features -> att2in2 ; all five losses (:449), backward, gradient all-reduce (N > 1), SGD.
    Host inputs are the image-side tensors only (C4 map, tokens, ROIs, one box + uint8 mask per expression): every
    target (response, box, mask) is built on the device."""

    KEYS = ("X", "labels", "e2i", "rois", "roi_labels", "gt_boxes", "gt_masks", "cap", "msk")

    def __init__(self, wl, device, world):
        from lang2seg_b200.nets.network import HotPathNet
        from lang2seg_b200.nets.res5_glue import Res5Glue
        from lang2seg_b200.parallel import FlatGradients
        torch.manual_seed(1234)
        self.net = HotPathNet(dict(seq_length=wl["L"], vocab_size=wl["V"], C4_feat_dim=wl["C"]),
                              head_to_tail=Res5Glue(wl["C"], 512, 3)).to(device).eval()
        self.params = [p for p in self.net.parameters() if p.requires_grad]
        self.opt = torch.optim.SGD(self.params, lr=1e-6, momentum=0.9, fused=True)
        self.flat = FlatGradients(self.net.gradient_groups()) if world > 1 else None

    def __call__(self, d, meta):
        net = self.net
        self.opt.zero_grad(set_to_none=True)
        loss = net.chained_train_step(d["X"], d["labels"], d["e2i"], d["rois"], d["roi_labels"], d["gt_boxes"],
                                      d["gt_masks"], d["cap"], d["msk"], meta["num_fg"], lengths=meta.get("lens"),
                                      steps=meta.get("steps"))
        loss.backward()
        if self.flat is not None:
            self.flat.pack()
            self.flat.all_reduce()
        self.opt.step()
        net._predictions.clear()
        net._losses.clear()
        net._proposal_targets = {}
        return loss.detach()


def build_branch_graphs(step, d, nocomm=False):
    """N > 1: one CUDA graph per branch (caption | mask head | encoder..crops), replayed concurrently on their own
    streams, + the encoder's backward + the SGD update.  Each gradient group's all-reduce is launched (outside the
    graphs, on NCCL's stream) behind ITS branch only, so it overlaps the other branches.  Returns (run, losses)."""
    def capture(fn, pool=None):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            out = fn()
        return g, out

    parts = step.parts
    flat = step.flat
    losses = []
    g_cap = g_mask = None
    if "caption" in parts:
        g_cap, l = capture(lambda: step.branch_caption(d))
        losses.append(l)
    if "mask" in parts:
        g_mask, l = capture(lambda: step.branch_mask(d))
        losses.append(l)
    g_main, l = capture(lambda: step.branch_main(d))
    losses.append(l)
    g_enc, _ = capture(step.bwd_rest, pool=g_main.pool())
    g_sgd, _ = capture(step.opt.step, pool=g_main.pool())
    s_cap = step.side if step.side is not None else torch.cuda.Stream(step.device)
    s_mask = step.side2 if step.side2 is not None else torch.cuda.Stream(step.device)
    ar = (lambda names: []) if nocomm else flat.all_reduce_async

    def run():
        cur = torch.cuda.current_stream()
        works = []
        if g_cap is not None:
            s_cap.wait_stream(cur)
            with torch.cuda.stream(s_cap):
                g_cap.replay()
                works += ar(["caption"])
        if g_mask is not None:
            s_mask.wait_stream(cur)
            with torch.cuda.stream(s_mask):
                g_mask.replay()
                works += ar(["heads"])
        g_main.replay()
        works += ar(["filter_generator"])
        g_enc.replay()
        works += ar(["encoder"])
        if g_cap is not None:
            cur.wait_stream(s_cap)
        if g_mask is not None:
            cur.wait_stream(s_mask)
        flat.wait(works)
        g_sgd.replay()
    return run, losses


def run_chained(wl, dev, world, rank, dist_on, steps=3, warm=2):
    """Device-timed and end-to-end (pinned host inputs every step) expressions/s of ChainedStep."""
    from lang2seg_b200 import synth
    from lang2seg_b200.pipeline import HostBatchPipeline
    E = wl["I"] * wl["EPI"]
    g = torch.Generator().manual_seed(4321 + rank)
    host = synth.chain_batch(g, wl["I"], wl["EPI"], wl["C"], wl["H"], wl["W"], wl["R"], wl["NFG"], wl["L"], wl["V"])
    meta = host.pop("_meta")
    step = ChainedStep(wl, dev, world)
    d = {k: v.to(dev) for k, v in host.items()}
    torch.backends.cudnn.benchmark = True       # as the reference does for its fixed-size TRAIN step (:585)
    for _ in range(warm):
        loss = step(d, meta)
    assert torch.isfinite(loss).all(), "chained step produced a non-finite loss"
    ms = time_region(lambda: step(d, meta), steps, dist_on) / steps
    pinned = {k: v.pin_memory() for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    pipe = HostBatchPipeline(dev, depth=2)
    out = []

    def e2e_run(n):
        pipe.submit(pinned)
        for i in range(n):
            if i + 1 < n:
                pipe.submit(pinned)
            dd = pipe.get()
            out.append(float(step(dd, meta)))
            pipe.release()

    ms_e2e = time_region(lambda: e2e_run(steps), 1, dist_on) / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    del step, d
    torch.cuda.empty_cache()
    return {"value": E * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warm,
            "e2e": {"value": E * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "res5": "resnet.layer4 (3 bottlenecks 1024->2048, stride 1, frozen BN folded) on cuDNN fp32, TF32 off -- glue, "
                    "not one of this repository's kernels (SURVEY 8 row a8)",
            "res5_rois_per_step": E * wl["R"], "res5_maps_per_step": wl["I"] + E,
            "launch": "eager (the chain is dominated by the cuDNN convolutions)", "peak_mem_gib": round(peak_gb, 1),
            "what": "dynfilter -> 7x7 crop -> res5 -> box head + mask head (fg) ; maps -> res5 -> caption features -> "
                    "att2in2 ; response / cls / box / mask / caption losses, backward, SGD; targets built on the device"}


def time_region(fn, steps, dist_on):
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t[0])
    return ms


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def component_rooflines(wl, d, step, pk):
    """Per-kernel device time (CUDA events, kernel alone => burst peaks) and roofline fractions.
    Algorithmic bytes / flops per SURVEY 8d (stated again in DESIGN.md)."""
    import lang2seg_b200.functional as F
    from lang2seg_b200 import _lib
    from lang2seg_b200._lib import call, ptr, stream
    I, EPI, C, H, W, Rn, NFG = (wl[k] for k in ("I", "EPI", "C", "H", "W", "R", "NFG"))
    parts = wl["parts"]
    E, HW = I * EPI, H * W
    dev = d["X"].device
    out = []
    filt = torch.tanh(torch.randn(E, 7, C, device=dev) * 0.1)
    fuse = torch.tanh(torch.randn(E, 7, device=dev))
    tgt = d.get("resp_tgt")
    # --- dynamic filter
    resp, Y, rl = F.dynamic_filter(d["X"], filt, fuse, d["e2i"], "sigmoid", tgt)
    rk = torch.empty(E, 7, H, W, device=dev)
    lossb = torch.empty(E, device=dev)
    nbf = _lib.size("l2s_dynfilter_fwd_workspace_bytes", I, E, C, H, W)
    wsf = torch.empty(nbf, dtype=torch.uint8, device=dev)
    t = ev_time(lambda: call("l2s_dynfilter_fwd", ptr(d["X"]), ptr(filt), ptr(fuse), ptr(d["e2i"]), ptr(resp), ptr(rk),
                             ptr(Y), ptr(tgt), ptr(lossb if tgt is not None else None), I, E, C, H, W, 0, ptr(wsf), nbf,
                             stream()))
    out.append(dict(kernel="dynfilter_fwd", ms=t, bound="hbm", work=4.0 * C * HW * (I + E)))
    if not step.fwd_only:
        dY = torch.randn_like(Y) * 1e-3
        dX, dfilt, dfuse = torch.empty_like(d["X"]), torch.empty_like(filt), torch.empty_like(fuse)
        gs = torch.ones(E, device=dev)
        nb = _lib.size("l2s_dynfilter_bwd_workspace_bytes", I, E, C, H, W)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        t = ev_time(lambda: call("l2s_dynfilter_bwd", ptr(d["X"]), ptr(filt), ptr(fuse), ptr(d["e2i"]), ptr(resp), ptr(rk),
                                 ptr(dY), None, ptr(tgt), ptr(gs if tgt is not None else None), ptr(dX), ptr(dfilt),
                                 ptr(dfuse), I, E, C, H, W, 0, ptr(ws), nb, stream()))
        out.append(dict(kernel="dynfilter_bwd", ms=t, bound="hbm", work=4.0 * C * HW * (E + 2 * I)))
        del dY, dX
    # --- ROI crop (7x7 and / or 14x14 samples + 2x2 max): 4*C*(49R+HW) B per expression and direction (+1 B/elt argmax)
    N = E * Rn
    for part, flags, tag in (("crop7", 0, "roi_crop"), ("crop_max", 1, "roi_crop_max")):
        if part not in parts:
            continue
        pool = torch.empty(N, C, 7, 7, device=dev)
        arg = torch.empty(N, C, 7, 7, device=dev, dtype=torch.uint8) if flags else None
        nb2 = _lib.size("l2s_roi_crop_workspace_bytes", E, N, flags)
        ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
        work = (4.0 + (1.0 if flags else 0.0)) * C * 49 * Rn * E + 4.0 * C * HW * E
        t = ev_time(lambda: call("l2s_roi_crop_fwd", ptr(Y), ptr(d["rois"]), ptr(pool), ptr(arg), E, C, H, W, N, 7, flags,
                                 0.0, 0.0, ptr(ws2), nb2, stream()))
        out.append(dict(kernel=tag + "_fwd", ms=t, bound="hbm", work=work))
        if not step.fwd_only:
            gp = step.g_pool_max if flags else step.g_pool
            dYb = torch.empty_like(Y)
            # as in the step: the backward reuses the ROI binning / geometry records the forward left in ws2
            t = ev_time(lambda: call("l2s_roi_crop_bwd", ptr(gp), ptr(d["rois"]), ptr(arg), ptr(dYb), E, C, H, W, N, 7,
                                     flags | F.CROP_WS_PREPARED, 0.0, 0.0, ptr(ws2), nb2, stream()))
            out.append(dict(kernel=tag + "_bwd", ms=t, bound="hbm", work=work))
            del dYb
        del pool, arg
    # --- Caffe-style RoI max-pool (POOLING_MODE != 'crop'; not part of the step, timed for its roofline only):
    # fwd reads <= the map once per expression + writes pool and argmax, bwd reads both and writes the map
    if "crop7" in parts and wl is WORKLOADS["cfg2"]:
        pool = torch.empty(N, C, 7, 7, device=dev)
        arg = torch.empty(N, C, 7, 7, device=dev, dtype=torch.int32)
        work = 8.0 * C * 49 * Rn * E + 4.0 * C * HW * E
        t = ev_time(lambda: call("l2s_roi_maxpool_fwd", 7, 7, 1.0 / 16, ptr(Y), ptr(d["rois"]), ptr(pool), ptr(arg), E, C, H, W, N,
                                 stream()))
        out.append(dict(kernel="roi_maxpool_fwd (not in the step)", ms=t, bound="hbm", work=work))
        if not step.fwd_only:
            dYb = torch.empty_like(Y)
            t = ev_time(lambda: call("l2s_roi_maxpool_bwd", 7, 7, 1.0 / 16, ptr(step.g_pool), ptr(d["rois"]), ptr(dYb), ptr(arg),
                                     E, C, H, W, N, stream()))
            out.append(dict(kernel="roi_maxpool_bwd (not in the step)", ms=t, bound="hbm", work=work))
            del dYb
        del pool, arg
    # --- mask head (tensor bound): fwd 2*n*49*2048*1024 + 2*n*196*256*81 ; bwd = 2x
    if "mask" in parts:
        n = E * NFG
        net = step.net
        with torch.no_grad():
            t = ev_time(lambda: F.mask_head(d["fc7"], net.mask_up_sampling.weight, net.mask_up_sampling.bias,
                                            net.mask_pred_net.weight, net.mask_pred_net.bias), iters=3, warm=1)
        flops_f = 2.0 * n * 49 * 2048 * 1024 + 2.0 * n * 196 * 256 * 81
        out.append(dict(kernel="mask_head_fwd (2 GEMM + repack)", ms=t, bound="tensor", work=flops_f))
        if not step.fwd_only:
            fc7 = d["fc7"].detach().requires_grad_(True)
            sc, p, ml = F.mask_head_with_loss(fc7, net.mask_up_sampling.weight, net.mask_up_sampling.bias,
                                              net.mask_pred_net.weight, net.mask_pred_net.bias, d["mlab"], d["mtgt"])
            t = ev_time(lambda: torch.autograd.grad(ml, [fc7, net.mask_up_sampling.weight, net.mask_pred_net.weight],
                                                    retain_graph=True), iters=3, warm=1)
            out.append(dict(kernel="mask_head_bwd (in-step: fused loss backward, dF + dWd GEMMs)", ms=t, bound="tensor",
                            work=2 * flops_f))
            del sc, p, ml
        # --- the dominant GEMM as the step runs it: GEMM1 of the mask head [49n x 2048] x [1024 x 2048]^T with the
        # bias + ReLU -> bf16-plane epilogue (EpiUp), timed through l2s_mask_head_gemm1 on the step's own operands
        M, Nn, K = n * 49, 1024, 2048
        g1 = F.mask_head_stage_runner(d["fc7"], net.mask_up_sampling.weight, net.mask_up_sampling.bias,
                                      net.mask_pred_net.weight, net.mask_pred_net.bias, stages=2)
        t = ev_time(g1, iters=5, warm=2)
        del g1
        out.append(dict(kernel="gemm_bf16x3_kernel<EpiUp> (mask head GEMM1, in-step epilogue)", ms=t, bound="tensor",
                        work=2.0 * M * Nn * K))
    # --- attention step (L2/HBM bound): one step fwd
    if "caption" in parts:
        A, Dd = 196, 512
        att_h = torch.randn(E, Dd, device=dev)
        feats = torch.randn(E, A, Dd, device=dev)
        p_att = torch.randn(E, A, Dd, device=dev)
        aw, ab = torch.randn(Dd, device=dev) * 0.04, torch.zeros(1, device=dev)
        wgt, res = torch.empty(E, A, device=dev), torch.empty(E, Dd, device=dev)
        t = ev_time(lambda: call("l2s_att_step_fwd", ptr(att_h), ptr(feats), ptr(p_att), ptr(aw), ptr(ab), ptr(wgt), ptr(res),
                                 E, A, Dd, Dd, stream()), iters=20)
        out.append(dict(kernel="att_step_fwd", ms=t, bound="hbm", work=2.0 * A * Dd * 4 * E))
    traffic = ncu_traffic() if wl is WORKLOADS["cfg2"] else {}
    for o in out:
        o["traffic"] = traffic.get(TRAFFIC_KEY.get(o["kernel"], ""))
        sec = o["ms"] * 1e-3
        if o["bound"] == "hbm":
            o["achieved"] = o["work"] / sec / 1e9
            o["peak"], o["unit"] = pk["hbm"], "GB/s"
        else:
            o["achieved"] = o["work"] / sec / 1e12
            o["peak"], o["unit"] = pk["tf_burst"], "TFLOP/s"
        o["frac"] = o["achieved"] / o["peak"]
    return out


def cpu_sample(wl, repeats=1, threads=None):
    """The oracle's torch-CPU port of the reference modules on ONE image and its expressions (the reference's
    native batching), the workload's parts, fwd+bwd (fwd only for cfg-1), timed on the host
    cores."""
    from oracle import restate as R
    from lang2seg_b200.layers.lang_encoder import RNNEncoder    # only as the container of default-initialised weights
    if threads:
        torch.set_num_threads(threads)
    w1 = dict(wl, I=1)
    parts, fwd_only = wl["parts"], bool(wl.get("fwd_only"))
    d = make_inputs(w1, 4321, "cpu")
    E, C, V, L = w1["EPI"], w1["C"], w1["V"], w1["L"]
    torch.manual_seed(7)
    enc = {k: v.detach().clone().requires_grad_(v.is_floating_point())
           for k, v in RNNEncoder(V, 512, 512, 512, bidirectional=True, n_layers=1).state_dict().items()}
    P = lambda *s: (torch.randn(*s) * 0.01).requires_grad_(True)      # noqa: E731
    dyn_w, dyn_b = [P(C, 1024) for _ in range(7)], [P(C) for _ in range(7)]
    rw, rb = P(7, 1024), P(7)
    up_w, up_b, pw, pb = P(2048, 256, 2, 2), P(256), P(81, 256, 1, 1), P(81)
    D = 512
    cp = {"att_embed.0.weight": P(D, 4096), "att_embed.0.bias": P(D), "ctx2att.weight": P(D, D), "ctx2att.bias": P(D),
          "embed.0.weight": P(V + 1, D), "logit.weight": P(V + 1, D), "logit.bias": P(V + 1),
          "core.i2h.weight": P(5 * D, D), "core.i2h.bias": P(5 * D), "core.h2h.weight": P(5 * D, D),
          "core.h2h.bias": P(5 * D), "core.a2c.weight": P(2 * D, D), "core.a2c.bias": P(2 * D),
          "core.attention.h2att.weight": P(D, D), "core.attention.h2att.bias": P(D),
          "core.attention.alpha_net.weight": P(1, D), "core.attention.alpha_net.bias": P(1)}
    g_pool = torch.randn(E * w1["R"], C, 7, 7) * 1e-4
    g_Y = torch.randn(E, C, w1["H"], w1["W"]) * 1e-4
    params = [v for v in enc.values() if v.requires_grad] + dyn_w + dyn_b + [rw, rb, up_w, up_b, pw, pb] + list(cp.values())

    def one():
        for p in params:
            p.grad = None
        with torch.set_grad_enabled(not fwd_only):
            X = d["X"].clone().requires_grad_(not fwd_only)
            _, hidden, _ = R.rnn_encoder_packed(d["labels"], enc)   # the reference's pack -> nn.LSTM -> unpack calls
            filt, fuse = R.filter_generator(hidden, dyn_w, dyn_b, rw, rb)
            r, Y = R.dynamic_filter(X, filt, fuse, d["e2i"].tolist())
            loss = torch.zeros(())
            outs, grads = [], []
            if "resp" in parts:
                loss = loss + R.response_loss(r, d["resp_tgt"]).sum()
            if "crop_max" in parts:
                outs.append(R.crop_pool(Y, d["rois"], max_pool=True)); grads.append(g_pool)
            if "crop7" in parts:
                outs.append(R.crop_pool(Y, d["rois"])); grads.append(g_pool)
            if "dY" in parts:
                outs.append(Y); grads.append(g_Y)
            if "mask" in parts:
                fc7 = d["fc7"].clone().requires_grad_(not fwd_only)
                s, _ = R.mask_head(fc7, up_w, up_b, pw, pb)
                loss = loss + R.mask_loss(s, d["mlab"], d["mtgt"])
            if "caption" in parts:
                att = d["att"].clone().requires_grad_(not fwd_only)
                loss = loss + R.caption_loss(d["fc"], att, d["cap"], d["msk"], cp)
            if not fwd_only:
                torch.autograd.backward([loss] + outs, [torch.ones(())] + grads)

    one()   # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        one()
    dt = (time.perf_counter() - t0) / repeats
    return E / dt, dt, E


def cpu_chain_sample(wl, threads=None):
    """The res5-chained step on the host cores: the reference's own modules when they are installed (kind "reference"),
    else the oracle port below."""
    if reference_modules_available():
        v, dt, e = cpu_sample_reference(wl, with_res5=True, threads=threads)
        return {"value": v, "unit": UNIT, "seconds_per_pass": dt, "kind": "reference",
                "sample": "1 image x %d expressions x %d ROIs, one fwd+bwd pass of the reference's own _predict chain with "
                          "resnet.layer4 (no warm-up pass)" % (e, wl["R"])}
    return _cpu_chain_sample_port(wl, threads)


def _cpu_chain_sample_port(wl, threads=None):
    """The res5-chained TRAIN step (oracle.restate.chained_train_losses: the reference's _predict + _add_losses from the
    dynamic filter on, resnet.layer4 included) on ONE image and its expressions, fwd+bwd on the host cores."""
    from oracle import restate as R
    from lang2seg_b200 import synth
    from lang2seg_b200.nets.network import HotPathNet          # only as the container of default-initialised weights
    from lang2seg_b200.nets.res5_glue import Res5Glue
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(4321)
    d = synth.chain_batch(g, 1, wl["EPI"], wl["C"], wl["H"], wl["W"], wl["R"], wl["NFG"], wl["L"], wl["V"])
    meta = d.pop("_meta")
    torch.manual_seed(7)
    net = HotPathNet(dict(seq_length=wl["L"], vocab_size=wl["V"], C4_feat_dim=wl["C"]), head_to_tail=Res5Glue(wl["C"], 512, 3))
    p = {(("res5." + k[len("_head."):]) if k.startswith("_head.") else k): v.detach().clone().requires_grad_(v.is_floating_point())
         for k, v in net.state_dict().items()}
    enc = {k[len("rnn_encoder."):]: v for k, v in p.items() if k.startswith("rnn_encoder.")}
    t0 = time.perf_counter()
    X = d["X"].clone().requires_grad_(True)
    _, hidden, _ = R.rnn_encoder_packed(d["labels"], enc)
    _, total = R.chained_train_losses(X, hidden, p, d["e2i"].tolist(), d["rois"], d["roi_labels"], d["gt_boxes"],
                                      d["gt_masks"].numpy(), d["cap"], d["msk"], meta["num_fg"])
    total.backward()
    dt = time.perf_counter() - t0
    E = wl["EPI"]
    return {"value": E / dt, "unit": UNIT, "seconds_per_pass": dt, "kind": "port",
            "sample": "1 image x %d expressions x %d ROIs, one fwd+bwd pass of the res5-chained step (no warm-up pass)" % (E, wl["R"])}


def reference_modules_available():
    """True when the reference's own python modules can be imported under oracle/shim.py: /root/reference in the
    development container, or the copy that oracle/install_reference.py leaves under baseline/_ref (travels to the box)."""
    try:
        from oracle import shim
        return shim.available()
    except Exception:
        return False


_REF_NET = {}


def cpu_sample_reference(wl, with_res5=False, threads=None, warm=False):
    """kind "reference": ONE image and its expressions through the reference's OWN unmodified modules (oracle/shim.py,
    BASELINE.md section 2 D1-D8), fwd+bwd, one expression at a time as the reference does:
    `RNNEncoder` + `Network._predict` of nets.resnet_v1_cycle_response.resnetv1 (dynamic filter, `_crop_pool_layer`,
    `_region_classification`, `_mask_prediction`; backbone and RPN stubbed: D6/D7) + `caption_models` att2in2 +
    `LanguageModelCriterion`; the loss terms are the reference's own expressions (network_cycle_response.py:404-422).
    Without res5 `_head_to_tail` is replaced by the synthetic features of the workload (as on the B200 arm) and the
    upstream gradient of pool5 stands in for res5's backward; with res5 the reference's `resnet.layer4` runs in the chain
    (crop -> layer4 -> heads ; maps -> layer4 -> caption features, :424-438)."""
    import torch.nn.functional as F
    from oracle import shim
    from lang2seg_b200 import synth
    if threads:
        torch.set_num_threads(threads)
    parts, fwd_only = wl["parts"], bool(wl.get("fwd_only"))
    import contextlib
    key = (wl["L"], wl["V"], wl["C"])
    if key not in _REF_NET:
        with contextlib.redirect_stdout(sys.stderr), shim.cpu_only():   # the reference prints while it builds: stdout is the ONE JSON line
            _REF_NET[key] = shim.build_reference_net(dict(seq_length=wl["L"], vocab_size=wl["V"], C4_feat_dim=wl["C"]), seed=7)
    net = _REF_NET[key]
    import misc.utils as ref_utils                    # the reference's lib/misc/utils.py (on sys.path after shim.install())
    crit = ref_utils.LanguageModelCriterion()
    w1 = dict(wl, I=1)
    E, Rn, NFG, H, W = w1["EPI"], w1["R"], w1["NFG"], w1["H"], w1["W"]
    if with_res5:
        g = torch.Generator().manual_seed(4321)
        d = synth.chain_batch(g, 1, E, w1["C"], H, W, Rn, NFG, w1["L"], w1["V"])
        meta = d.pop("_meta")
    else:
        d = make_inputs(w1, 4321, "cpu")
    g_pool = torch.randn(Rn, w1["C"], 7, 7) * 1e-4
    g_Y = torch.randn(1, w1["C"], H, W) * 1e-4
    orig_head = type(net)._head_to_tail

    def one():
        net.zero_grad()
        with torch.set_grad_enabled(not fwd_only):
            total = torch.zeros(())
            for e in range(E):                          # the reference runs one (image, expression) pair per step
                X = d["X"][:1].clone().requires_grad_(not fwd_only)
                labels = d["labels"][e:e + 1]
                labels = labels[:, :max(1, int((labels != 0).sum()))]     # as Network.forward does (:634-636)
                if with_res5:
                    nfg = NFG
                    rois_e = torch.cat([d["rois"][e * NFG:(e + 1) * NFG], d["rois"][E * NFG + e * (Rn - NFG):E * NFG + (e + 1) * (Rn - NFG)]]).clone()
                    rois_e[:, 0] = 0
                    net._head_to_tail = lambda pool5: orig_head(net, pool5)
                    stash = {}
                else:
                    nfg = NFG if "mask" in parts else 0
                    rois_e = d["rois"][e * Rn:(e + 1) * Rn].clone() if Rn else torch.zeros(1, 5)
                    rois_e[:, 0] = 0
                    fc7 = d["fc7"][e * NFG:(e + 1) * NFG].clone().requires_grad_(not fwd_only) if "mask" in parts else None
                    stash = {}

                    def head(pool5, fc7=fc7, stash=stash):
                        stash["pool5"] = pool5
                        out = pool5.new_zeros(pool5.shape[0], 2048, 7, 7)
                        if fc7 is not None:
                            out = torch.cat([fc7, out[fc7.shape[0]:]], 0)
                        return out
                    net._head_to_tail = head
                gated, _, _, _, _ = shim.run_predict(net, X, labels, rois_e, mode="TRAIN", num_fg=max(nfg, 1) if not with_res5 and nfg == 0 else nfg)
                loss = torch.zeros(())
                if "resp" in parts or with_res5:
                    resp = net._predictions["response"].squeeze(1).squeeze(0)
                    tgt = d["resp_tgt"][e] if not with_res5 else torch.from_numpy(
                        __import__("scipy.misc", fromlist=["imresize"]).imresize(d["gt_masks"][e].numpy(), (H, W), interp="nearest").astype("float32"))
                    loss = loss + F.binary_cross_entropy_with_logits(resp, tgt)                      # :415-422
                if with_res5 or "mask" in parts:
                    ms = net._predictions["mask_score"]
                    lab = (d["roi_labels"][e * NFG:(e + 1) * NFG].long() if with_res5 else d["mlab"][e * NFG:(e + 1) * NFG])
                    idx = lab.view(-1, 1, 1, 1).expand(lab.numel(), 1, 14, 14)
                    mt = (torch.zeros(lab.numel(), 14, 14) if with_res5 else d["mtgt"][e * NFG:(e + 1) * NFG])
                    loss = loss + F.binary_cross_entropy_with_logits(torch.gather(ms, 1, idx).squeeze(1), mt)   # :404-413
                if with_res5:
                    cls = net._predictions["cls_score"]
                    rl = torch.cat([d["roi_labels"][e * NFG:(e + 1) * NFG], torch.zeros(Rn - NFG)]).long()
                    loss = loss + F.cross_entropy(cls.view(-1, cls.shape[-1]), rl)                    # :390-394
                    fb = orig_head(net, net._predictions["net_conv_before"])
                    fa = orig_head(net, gated)
                    fcf = torch.cat((fb.mean(3).mean(2), fa.mean(3).mean(2)), 1)
                    attf = torch.cat((F.adaptive_avg_pool2d(fb, [14, 14]).permute(0, 2, 3, 1).contiguous(),
                                      F.adaptive_avg_pool2d(fa, [14, 14]).permute(0, 2, 3, 1).contiguous()), 3)   # :424-438
                    cap, msk = d["cap"][e:e + 1], d["msk"][e:e + 1]
                    loss = loss + crit(net.caption_model(fcf, attf, cap), cap[:, 1:], msk[:, 1:])
                else:
                    if "crop7" in parts or "crop_max" in parts:
                        loss = loss + (stash["pool5"] * g_pool).sum()
                    if "dY" in parts:
                        loss = loss + (gated * g_Y).sum()
                    if "caption" in parts:
                        att = d["att"][e:e + 1].clone().requires_grad_(not fwd_only)
                        cap, msk = d["cap"][e:e + 1], d["msk"][e:e + 1]
                        loss = loss + crit(net.caption_model(d["fc"][e:e + 1], att, cap), cap[:, 1:], msk[:, 1:])
                total = total + loss
            if not fwd_only:
                total.backward()
        net._head_to_tail = lambda pool5: orig_head(net, pool5)

    with contextlib.redirect_stdout(sys.stderr), shim.cpu_only():      # the reference's .cuda() calls stay on the host
        if warm:
            one()
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
    return E / dt, dt, E


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use_ref = reference_modules_available()
    sample_fn = cpu_sample_reference if use_ref else (lambda w: cpu_sample(w, 1))
    kind = "reference" if use_ref else "port"
    vals = []
    if use_ref:
        try:
            cpu_sample_reference(wl, warm=True)            # build the net + one untimed pass even with --warmup 0
        except Exception as exc:                           # fall back to the oracle port rather than print nothing
            print("bench: reference modules failed (%s); timing the oracle port" % str(exc).splitlines()[0], file=sys.stderr)
            use_ref = False
            sample_fn, kind = (lambda w: cpu_sample(w, 1)), "port"
    for _ in range(args.warmup):
        sample_fn(wl)
    n_expr = 0
    for _ in range(args.steps):
        v, dt, e = sample_fn(wl)
        n_expr += e
        vals.append(dt)
    total = sum(vals)
    value = n_expr / total
    sample = ("1 image x %d expressions of the workload per step (reference's native batching), fwd+bwd, torch CPU; " % wl["EPI"]) + \
             ("the reference's own unmodified modules under oracle/shim.py (baseline/_ref)" if use_ref else
              "oracle torch-CPU port (the reference modules are not installed under baseline/_ref)")
    chained = None
    if not args.no_res5 and all(p in wl["parts"] for p in ("resp", "crop7", "mask", "caption")):
        try:
            chained = cpu_chain_sample(wl)
        except Exception as exc:
            chained = {"unavailable": str(exc).splitlines()[0][:200]}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu": cpu_model(), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "with_res5": chained, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (overrides the workload's I; cfg5 sweeps 8..256 total)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-components", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    ap.add_argument("--no-res5", action="store_true",
                    help="skip the second, res5-chained number (reference-faithful chain with resnet.layer4 as cuDNN glue)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.batch > 0:
        wl = dict(wl, I=args.batch, name=wl["name"] + " [--batch %d images per GPU]" % args.batch)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch.distributed as dist
    from lang2seg_b200 import _lib
    _lib.load()          # fails loudly if libl2s.so is missing: there is no fallback
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        # keep stdout to the ONE JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pk = peaks()
    E = wl["I"] * wl["EPI"]

    force_split = os.environ.get("L2S_BENCH_FORCE_SPLIT") == "1"      # diagnostics: the N > 1 graph structure on one GPU
    step = HotPathStep(wl, dev, 2 if (force_split and world == 1) else world)
    d = make_inputs(wl, 1234 + rank, dev)
    for _ in range(args.warmup):
        step(d)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    step(d)
    launches_per_step = _lib.launch_count() - l0          # kernels of libl2s.so per step (counted on an eager step)
    # The step (shapes static for a given batch geometry) is captured in CUDA graphs and replayed: the timed region then
    # measures the kernels, not Python / launch latency.
    #   N = 1: one graph for the whole step (lang encoder .. SGD update).
    #   N > 1: three graphs with the NCCL all-reduces launched between them, so that communication overlaps compute:
    #          A = forward + the backward of all three branches down to the expression embedding
    #              -> all-reduce(caption), all-reduce(heads), all-reduce(filter_generator) start
    #          B = the rest of the backward (language encoder), concurrent with them -> all-reduce(encoder)
    #          C = fused SGD update over the flat gradient views.
    whole = (world == 1 and not force_split) or step.fwd_only
    run, graphed, launch_desc = (lambda: step(d)), False, "eager launches"
    if not args.no_graph:
        def capture(fn, pool=None):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                out = fn()
            return g, out

        def build_graphs():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(d)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if whole:
                graph, static_loss = capture(lambda: step(d))
                return graph.replay, static_loss
            g_a, static_loss = capture(lambda: step.fwd_bwd(d, split=True))
            g_b, _ = capture(step.bwd_rest, pool=g_a.pool())
            g_c, _ = capture(step.opt.step, pool=g_a.pool())
            early = [n for n in step.EARLY_GROUPS]
            late = [n for n in step.flat.names if n not in early]
            nocomm = os.environ.get("L2S_BENCH_NOCOMM") == "1"    # diagnostics: the three graphs without the collectives

            def run_split():
                g_a.replay()
                w = [] if nocomm else step.flat.all_reduce_async(early)
                g_b.replay()
                w += [] if nocomm else step.flat.all_reduce_async(late)
                step.flat.wait(w)
                g_c.replay()
            return run_split, static_loss

        def all_ranks_ok(ok):
            """a capture failure on one rank must take every rank down the same path (the collectives have to match)"""
            if not dist_on:
                return ok
            flag = torch.tensor([0 if ok else 1], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            return int(flag) == 0

        def build_branch():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(d)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            run_b, losses = build_branch_graphs(step, d, nocomm=os.environ.get("L2S_BENCH_NOCOMM") == "1")
            run_b()
            return run_b, torch.stack([l.float() for l in losses]).sum()

        def one_stream():
            step.side = step.side2 = None

        # capture is an optimisation: structures are tried in this order, then eager launches (said so in the line)
        if whole:
            plans = [("one CUDA graph replay per step", None, build_graphs),
                     ("one CUDA graph replay per step", one_stream, build_graphs)]
        else:
            three = ("three CUDA graphs per step (fwd + backward down to the expression embedding | encoder backward | SGD) "
                     "with the NCCL all-reduces of the flat gradient groups launched between them")
            plans = [("one CUDA graph per branch (att2in2 | mask head | encoder..crops + encoder backward), replayed "
                      "concurrently on three streams, each gradient group all-reduced (NCCL) behind its own branch; SGD graph",
                      None, build_branch),
                     (three, None, build_graphs), (three, one_stream, build_graphs)]
            if os.environ.get("L2S_BENCH_BRANCH_GRAPHS") == "0":
                plans = plans[1:]
        for attempt, (desc, prep, builder) in enumerate(plans):
            err = None
            try:
                if prep is not None:
                    prep()
                if attempt == 0 and os.environ.get("L2S_BENCH_TEST_CAPTURE_FAIL") == "1":    # diagnostics: exercise the retry
                    raise RuntimeError("injected capture failure")
                run_g, static_loss = builder()
                run_g()
                torch.cuda.synchronize()
                assert torch.isfinite(static_loss).all()
            except Exception as exc:
                err = str(exc).splitlines()[0] if str(exc) else repr(exc)
            if all_ranks_ok(err is None):
                run, graphed, launch_desc = run_g, True, desc
                if rank == 0:
                    print("bench: loss of the graphed step %.9g (streams: %d; %s)" % (
                        float(static_loss), 1 + (step.side is not None) + (step.side2 is not None), desc[:40]), file=sys.stderr)
                break
            torch.cuda.synchronize()
            step._pending = None
            step.net._predictions.clear()
            step.net._losses.clear()
            more = attempt + 1 < len(plans)
            print("bench: CUDA graph capture failed (%s); %s" % (err, "trying the next structure" if more else
                                                                 "timing eager launches"), file=sys.stderr)
            if not more:
                run = lambda: step(d)     # noqa: E731
    streams_used = (1 + (step.side is not None) + (step.side2 is not None)) if graphed else 1    # parallel branches of the graph
    for _ in range(args.warmup):
        run()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = time_region(run, args.steps, dist_on)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    value = E * world * args.steps / (ms * 1e-3)

    # ---- e2e: same step through the public modules with HOST (pinned) inputs.  Every step copies its own inputs
    # host->device (lang2seg_b200.pipeline: copy stream + double buffer, so the copy of step i+1 overlaps the kernels
    # of step i) and reads its loss back (device->host) inside the timed region.
    from lang2seg_b200.pipeline import HostBatchPipeline
    host = make_inputs(wl, 1234 + rank, dev, pinned=True)
    host_meta = host.pop("_meta")
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    losses = []
    pipe = HostBatchPipeline(dev, depth=2)

    def e2e_run(n):
        pipe.submit(host)
        for i in range(n):
            if i + 1 < n:
                pipe.submit(host)
            dd = pipe.get()
            losses.append(float(step(dd, host_meta).detach()))     # device->host read of the step's result
            pipe.release()

    e2e_run(2)
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = time_region(lambda: e2e_run(e2e_steps), 1, dist_on)
    e2e_value = E * world * e2e_steps / (ms_e2e * 1e-3)
    assert pipe.bytes_per_batch == h2d

    comps = None
    if rank == 0 and not args.no_components:
        comps = component_rooflines(wl, d, step, pk)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU leg is timed at N = 1 only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref_ok = reference_modules_available()
        if ref_ok:
            try:
                reps = max(1, int(12.0 / max(1.0, 0.7 * wl["EPI"])))
                cpu_sample_reference(wl, warm=True)            # builds the reference net, warms up
                t0 = time.perf_counter()
                for _ in range(reps):
                    v, dt, e = cpu_sample_reference(wl)
                dt = (time.perf_counter() - t0) / reps
                v = e / dt
                kind, how = "reference", "the reference's own unmodified modules under oracle/shim.py (baseline/_ref)"
            except Exception as exc:      # the B200 line must survive a failure of the baseline leg: fall back to the port
                print("bench: reference modules failed (%s); timing the oracle port" % str(exc).splitlines()[0], file=sys.stderr)
                ref_ok = False
        if not ref_ok:
            v, dt, e = cpu_sample(wl, repeats=max(1, int(12.0 / max(1.0, 0.7 * wl["EPI"]))))
            kind, how = "port", "oracle torch-CPU port"
        cpu = {"value": v, "unit": UNIT, "cores": cores, "cpu": cpu_model(), "kind": kind,
               "sample": "1 image x %d expressions of the workload (reference's native batching), fwd+bwd, "
                         "%s, %.2f s per pass" % (e, how, dt)}
        if not args.no_res5 and all(p in wl["parts"] for p in ("resp", "crop7", "mask", "caption")):
            try:
                cpu["with_res5"] = cpu_chain_sample(wl)
            except Exception as exc:
                cpu["with_res5"] = {"unavailable": str(exc).splitlines()[0][:200]}
    # ---- the second, labelled number: the same path chained through res5 (cuDNN glue) like the reference's _predict
    chained = None
    if not args.no_res5 and all(p in wl["parts"] for p in ("resp", "crop7", "mask", "caption")) and not step.fwd_only:
        del step, d, host, pipe
        run = None
        torch.cuda.empty_cache()
        try:
            chained = run_chained(wl, dev, world, rank, dist_on)
        except Exception as exc:      # the graded line must survive a failure of the glue leg
            chained = {"unavailable": str(exc).splitlines()[0][:200]}
    if rank == 0:
        roof = None
        if comps:
            gemms = [c for c in comps if c["kernel"].startswith("gemm_")]
            dom = max((c for c in comps if not c["kernel"].startswith("gemm_") and "not in the step" not in c["kernel"]),
                      key=lambda c: c["ms"])
            if dom["bound"] == "tensor" and gemms:
                dom = gemms[0]
            roof = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": dom.get("traffic"), "peak_source": pk["source"],
                    "ms_per_launch": dom["ms"]}
            if dom["bound"] == "tensor":
                # fp32 parity (1e-4) needs the bf16x3 split product: 3 tensor passes per algorithmic flop, so `frac`
                # against the bf16 peak tops out at 1/3 ("fp32-exact variant: divide peak by 3")
                roof["tensor_passes_per_flop"] = 3
                roof["frac_of_fp32_exact_ceiling"] = 3.0 * dom["frac"]
                roof["traffic_note"] = ("DRAM bytes per launch of the in-step GEMM of this shape (EpiUp: bf16 plane "
                                        "output) from the ncu --set full capture, profiles/r01_traffic.json")
        part_names = {"resp": "response BCE", "crop7": "7x7 crop", "crop_max": "14x14+2x2max crop", "mask": "mask head",
                      "caption": "att2in2", "dY": "synthetic upstream gradient on the gated map"}
        includes = "lang encoder, filter generator, dynamic filter, " + ", ".join(part_names[p] for p in wl["parts"]) + \
                   (" -- forward only" if wl.get("fwd_only") else " -- fwd+bwd, grad all-reduce, SGD")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": wl["name"], "per_gpu_expressions": E, "parallelism": "dp%d" % world,
                           "l2": "working set per step (feature maps, ROI crops, res5 features) far larger than the 126 MB L2; no flush needed"
                           if "caption" not in wl["parts"] or wl["I"] * wl["EPI"] >= 32 else
                           "inputs + activations per step exceed the 126 MB L2 (att/fc features alone: %d MB)" % (wl["I"] * wl["EPI"] * 196 * 4096 * 4 // 2**20),
                           "includes": includes,
                           "launch": launch_desc,
                           "streams": streams_used},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                        "how": "pinned host inputs -> copy stream (double buffered, overlaps the previous step) -> step -> loss.item()"},
                "roofline": roof, "cpu_baseline": cpu, "with_res5": chained,
                "components": [{k: (round(v, 5) if isinstance(v, float) else v) for k, v in c.items()} for c in comps]
                if comps else None}
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

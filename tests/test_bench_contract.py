"""Host-side checks of the benchmark / packaging contract that need no GPU."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.module


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under lang2seg_b200/ may import it (no CPU fallback behind the API)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lang2seg_b200")):
        for f in files:
            if f.endswith(".py"):
                path = os.path.join(dirpath, f)
                for name in _imports(ast.parse(open(path).read())):
                    if name == "oracle" or name.startswith("oracle."):
                        bad.append((path, name))
    assert not bad, bad


def test_bench_uses_the_oracle_only_in_its_cpu_legs():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_sample", "cpu_sample_reference", "reference_modules_available", "cpu_chain_sample", "_cpu_chain_sample_port",
               "run_reference_arm"}
    for node in tree.body:
        names = [n for n in _imports(node) if n == "oracle" or n.startswith("oracle.")] if not isinstance(
            node, (ast.FunctionDef, ast.ClassDef)) else []
        assert not names, "module-level oracle import in bench.py"
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name not in allowed:
            assert not [n for n in _imports(node) if n == "oracle" or n.startswith("oracle.")], node.name


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) on one bounded sample."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "tiny"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "expressions/s" and d["higher_is_better"] is True
    # "reference" when the reference modules are installed (/root/reference or baseline/_ref), else the oracle port
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_recorded_bench_line_has_every_contract_key():
    """The line the B200 arm printed at HEAD (profiles/r02s3_bench_cfg2.json) carries the keys of the benchmark contract."""
    path = os.path.join(ROOT, "profiles", "r02s3_bench_cfg2.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("cfg2") and "model" not in d["config"]
    assert abs(d["value"] - 48 * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6            # 48 expressions per step
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 1e9 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and d["gpu_launches"] > 0
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_branch_streams_and_split_backward_leave_the_step_unchanged(monkeypatch):
    """bench.HotPathStep under CUDA-graph capture: (a) the three branch streams, (b) the N > 1 structure (backward cut at
    the generated filters, two graphs) must give the loss and the parameter gradients of the one-stream, one-graph step."""
    import torch
    sys.path.insert(0, ROOT)
    import bench

    wl = bench.WORKLOADS["tiny"]
    dev = torch.device("cuda:0")

    def run(streams, split):
        monkeypatch.setenv("L2S_BENCH_STREAMS", str(streams))
        step = bench.HotPathStep(wl, dev, 2 if split else 1)
        d = bench.make_inputs(wl, 1234, dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up outside capture (lazy initialisations)
            if split:
                step.fwd_bwd(d, split=True); step.bwd_rest()
            else:
                step.fwd_bwd(d)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            loss = step.fwd_bwd(d, split=split)
        graphs = [g1]
        if split:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                step.bwd_rest()
            graphs.append(g2)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in step.net.named_parameters() if p.grad is not None}
        return float(loss), grads, (step.side is not None) + (step.side2 is not None)

    def run_branch_graphs():
        """the N > 1 structure of bench.py: one graph per branch, replayed concurrently on three streams"""
        monkeypatch.setenv("L2S_BENCH_STREAMS", "2")
        step = bench.HotPathStep(wl, dev, 2)
        d = bench.make_inputs(wl, 1234, dev)
        p0 = {n: p.detach().clone() for n, p in step.net.named_parameters()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step.fwd_bwd(d, split=True); step.bwd_rest(); step.update()     # the optimizer's momentum state exists before capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                                               # ... and the parameters are those of the other runs
            for n, p in step.net.named_parameters():
                p.copy_(p0[n])
        run_b, losses = bench.build_branch_graphs(step, d)
        run_b()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in step.net.named_parameters() if p.grad is not None}
        moved = [n for n, p in step.net.named_parameters() if p.grad is not None and not torch.equal(p.detach(), p0[n])]
        return float(torch.stack([l.float() for l in losses]).sum()), grads, moved

    l0, g0, n0 = run(0, False)
    l1, g1_, n1 = run(2, False)
    l2, g2_, n2 = run(2, True)
    l3, g3_, moved = run_branch_graphs()
    assert n0 == 0 and n1 == 2 and n2 == 2
    assert l0 == l1 == l2
    assert abs(l3 - l0) <= 1e-6 * abs(l0)                                 # the three losses are added in another order
    assert set(g0) == set(g1_) == set(g2_) == set(g3_) and len(g0) > 20
    assert len(moved) > 20                                                # the SGD graph ran on the packed gradients
    for k in g0:
        den = float(g0[k].abs().max()) + 1e-30
        assert float((g0[k] - g1_[k]).abs().max()) / den < 1e-5, k      # split-K atomics: not bit-identical run to run
        assert float((g0[k] - g2_[k]).abs().max()) / den < 1e-5, k
        assert float((g0[k] - g3_[k]).abs().max()) / den < 1e-5, k

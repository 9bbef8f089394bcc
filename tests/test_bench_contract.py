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

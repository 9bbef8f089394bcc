#!/bin/bash
# the default bench line (cpu baseline + res5 chain) at HEAD, and the reference arm
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg2_s3.json 2> gpurun_out/r02_bench_cfg2_s3.err
echo "bench cfg2 (default) exit=$?" | tee -a gpurun_out/summary_final3.txt; tail -c 300 gpurun_out/r02_bench_cfg2_s3.err
python scripts/show_bench.py gpurun_out/r02_bench_cfg2_s3.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_cfg2_s3.json").read().strip().splitlines()[-1])
print("roofline", json.dumps(d["roofline"]))
print("cpu", json.dumps(d.get("cpu_baseline"))[:400])
print("res5", json.dumps(d.get("with_res5"))[:400])
print("config", json.dumps(d["config"])[:600])
PY

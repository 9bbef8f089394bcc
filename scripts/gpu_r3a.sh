#!/bin/bash
# CPU arm on the reference's own modules (baseline/_ref): reference arm line + default bench line
mkdir -p gpurun_out
ls baseline/_ref | head -3
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cfg2_refmods.json 2> gpurun_out/r02_bench_reference_cfg2_refmods.err
echo "reference arm exit=$?"; tail -c 300 gpurun_out/r02_bench_reference_cfg2_refmods.err | tail -n 2; cut -c1-900 gpurun_out/r02_bench_reference_cfg2_refmods.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg2_refmods.json 2> gpurun_out/r02_bench_cfg2_refmods.err
echo "bench cfg2 (default) exit=$?"; tail -c 300 gpurun_out/r02_bench_cfg2_refmods.err | tail -n 2
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_cfg2_refmods.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("cpu", json.dumps(d.get("cpu_baseline")))
PY
for w in cfg1 cfg4; do
  timeout 900 python bench.py --impl reference --workload $w --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_${w}_refmods.json 2> gpurun_out/r02_bench_reference_${w}_refmods.err
  echo "reference arm $w exit=$?"; cut -c1-300 gpurun_out/r02_bench_reference_${w}_refmods.json
done

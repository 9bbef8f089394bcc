#!/bin/bash
# round 2, visit A: full-size crop parity + bench lines of every BASELINE.json config with the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_roi.log 2>&1
echo "test_gpu_roi exit=$?" | tee -a gpurun_out/summary_r2a.txt
tail -n 15 gpurun_out/test_roi.log
for w in cfg2 cfg1 cfg3 cfg4 cfg5; do
  extra="--no-cpu-baseline"; [ $w = cfg2 ] && extra=""
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 $extra > gpurun_out/bench_${w}_r2a.json 2> gpurun_out/bench_${w}_r2a.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2a.txt; tail -c 600 gpurun_out/bench_${w}_r2a.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${w}_r2a.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"], 1), "expr/s", round(d["ms_per_step"], 3), "ms/step e2e", round(d["e2e"]["value"], 1))
    for c in d.get("components") or []:
        print("   %-70s %8.3f ms  frac %.3f" % (c["kernel"], c["ms"], c["frac"]))
except Exception as e:
    print("$w: no line", e)
PY
done

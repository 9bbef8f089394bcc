#!/bin/bash
# quick GPU visit: the parity suites named in $1 (default all), then the bench without the CPU baseline
mkdir -p gpurun_out
for t in ${1:-roi dynfilter att mask_head targets nms net}; do
  timeout 420 python -m pytest tests/test_gpu_$t.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit=$?" | tee -a gpurun_out/summary_quick.txt
  tail -n 3 gpurun_out/test_$t.log
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit=$?"; tail -c 400 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("ms/step %.3f value %.1f e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d["clocks"])
for c in d["components"]:
    print("   %-62s %8.4f ms  frac %.3f" % (c["kernel"][:62], c["ms"], c["frac"]))
PY

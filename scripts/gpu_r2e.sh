#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_att.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_att.log 2>&1
echo "test_gpu_att exit=$?" | tee -a gpurun_out/summary_r2e.txt
tail -n 12 gpurun_out/test_att.log
timeout 300 python scripts/prof_decode_phases.py --B 48 --L 10 > gpurun_out/decode_phases_B48.txt 2>&1; grep -E "==|step  [1-3]:|step 1[89]:" gpurun_out/decode_phases_B48.txt
timeout 300 python scripts/prof_decode_phases.py --B 16 --L 20 > gpurun_out/decode_phases_B16.txt 2>&1; grep -E "==|step  [1-3]:|step 1[89]:" gpurun_out/decode_phases_B16.txt
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_${w}_r2e.json 2> gpurun_out/bench_${w}_r2e.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2e.txt; tail -c 400 gpurun_out/bench_${w}_r2e.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2e.json
done

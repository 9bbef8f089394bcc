#!/bin/bash
# round 2, visit B: generic row-owner crop backward (max-pool / large maps)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_roi.log 2>&1
echo "test_gpu_roi exit=$?" | tee -a gpurun_out/summary_r2b.txt
tail -n 15 gpurun_out/test_roi.log
for w in cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2b.json 2> gpurun_out/bench_${w}_r2b.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2b.txt; tail -c 600 gpurun_out/bench_${w}_r2b.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2b.json
done

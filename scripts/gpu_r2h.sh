#!/bin/bash
# A/B of the barrier wait policy in the row-owner crop backward kernels (components of cfg2 / cfg3 / cfg5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_roi.log 2>&1
echo "test_gpu_roi exit=$?" | tee -a gpurun_out/summary_r2h.txt; tail -n 4 gpurun_out/test_roi.log
for mode in backoff parked; do
  for w in cfg2 cfg3 cfg5; do
    if [ $mode = parked ]; then export L2S_CROP_WAIT_PARKED=1; else unset L2S_CROP_WAIT_PARKED; fi
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2h_$mode.json 2> gpurun_out/bench_${w}_r2h_$mode.err
    echo "bench $w $mode exit=$?" | tee -a gpurun_out/summary_r2h.txt
    python scripts/show_bench.py gpurun_out/bench_${w}_r2h_$mode.json | grep -E "expr/s|roi_crop"
  done
done

#!/bin/bash
# round 2, visit G: row-mirror forward crop kernel + per-pooled-row backward; A/B of the generic kernels on the 7x7 path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_roi.log 2>&1
echo "test_gpu_roi exit=$?" | tee -a gpurun_out/summary_r2g.txt; tail -n 6 gpurun_out/test_roi.log
L2S_CROP_FWD_ROWS=1 L2S_CROP_BWD_ROWSX=1 timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_roi_generic.log 2>&1
echo "test_gpu_roi (generic kernels on 7x7) exit=$?" | tee -a gpurun_out/summary_r2g.txt; tail -n 6 gpurun_out/test_roi_generic.log
for w in cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2g.json 2> gpurun_out/bench_${w}_r2g.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2g.txt; tail -c 300 gpurun_out/bench_${w}_r2g.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2g.json
done
L2S_CROP_FWD_ROWS=1 L2S_CROP_BWD_ROWSX=1 timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_r2g_generic.json 2> gpurun_out/bench_cfg2_r2g_generic.err
echo "bench cfg2 generic exit=$?" | tee -a gpurun_out/summary_r2g.txt
python scripts/show_bench.py gpurun_out/bench_cfg2_r2g_generic.json

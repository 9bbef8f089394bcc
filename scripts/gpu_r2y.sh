#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mask_head.py -x -q -m gpu -p no:cacheprovider -k "gemm_f32 or wgrad" > gpurun_out/test_r2y.log 2>&1
echo "tests exit=$?"; tail -n 3 gpurun_out/test_r2y.log
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r2y.json 2> gpurun_out/bench_${w}_r2y.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2y.json | grep -E "expr/s"
done
timeout 600 python scripts/prof_step.py --workload cfg4 --steps 3 > gpurun_out/step_kernels_cfg4_r2y.txt 2>&1
grep -E "gemm_f32|sgemm|un-profiled|kernel time" gpurun_out/step_kernels_cfg4_r2y.txt | cut -c1-200

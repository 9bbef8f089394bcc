#!/bin/bash
# call g: embedding backward v3 (4 rows x 4 columns in flight per lane, 4 warps over row blocks), long-input colsum
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_att.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4g.log 2>&1
echo "pytest exit=$?"; tail -n 3 gpurun_out/test_r4g.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4g.json 2> gpurun_out/bench_cfg2_r4g.err
echo "bench cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_r4g.json 2>/dev/null | head -1
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 --trace embedding,colsum,reduce_kernel,att_accum > gpurun_out/step_kernels_cfg2_r4g.txt 2>&1
echo "prof_step exit=$?"; awk '/# timeline/{f=1} f' gpurun_out/step_kernels_cfg2_r4g.txt | cut -c1-150 | head -40

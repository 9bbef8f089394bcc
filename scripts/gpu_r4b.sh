#!/bin/bash
# call b: merged-column 7x7 row-owner backward (one RMW per distinct map column), odd-C dynamic filter fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dynfilter.py tests/test_gpu_roi.py -q -m gpu -p no:cacheprovider > gpurun_out/test_r4b.log 2>&1
echo "pytest exit=$?"; tail -n 5 gpurun_out/test_r4b.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg2_r4b.json 2> gpurun_out/bench_cfg2_r4b.err
echo "bench cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_r4b.json | head -6

#!/bin/bash
# warp-per-ROI RoI max-pool forward: parity + timing
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2r.txt
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider -k "maxpool" > gpurun_out/test_r2r.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2r.txt; tail -n 12 gpurun_out/test_r2r.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg2_r2r.json 2> gpurun_out/bench_cfg2_r2r.err
echo "bench cfg2 exit=$?" | tee -a gpurun_out/summary_r2r.txt; tail -c 300 gpurun_out/bench_cfg2_r2r.err
python scripts/show_bench.py gpurun_out/bench_cfg2_r2r.json | grep -E "expr/s|dynfilter|maxpool"

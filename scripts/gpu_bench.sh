#!/bin/bash
# bench + ncu launch list on one GPU
mkdir -p gpurun_out
timeout 300 python bench.py --workload tiny --steps 3 --warmup 3 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
echo "bench tiny exit=$?" | tee -a gpurun_out/summary.txt; tail -c 600 gpurun_out/bench_tiny.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
echo "bench cfg2 exit=$?" | tee -a gpurun_out/summary.txt; tail -c 1500 gpurun_out/bench_cfg2.err
cat gpurun_out/bench_cfg2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit=$?" | tee -a gpurun_out/summary.txt

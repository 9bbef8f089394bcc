#!/bin/bash
# round 2, visit D: full parity CI after the hygiene pass + launch lists (cfg2, cfg4) to time the persistent kernels
mkdir -p gpurun_out
bash scripts/gpu_ci.sh
for w in cfg2 cfg4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${w}_r2d.csv \
     python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/ncu_bench_${w}_r2d.log 2>&1
  echo "ncu launches $w exit=$?" | tee -a gpurun_out/summary.txt
  python scripts/summarize_launches.py gpurun_out/launches_${w}_r2d.csv > gpurun_out/launches_${w}_r2d_summary.txt
  head -45 gpurun_out/launches_${w}_r2d_summary.txt
done
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_${w}_r2d.json 2> gpurun_out/bench_${w}_r2d.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary.txt; tail -c 400 gpurun_out/bench_${w}_r2d.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2d.json
done

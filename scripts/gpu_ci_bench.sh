#!/bin/bash
# full GPU parity CI + default bench lines of the given workloads: bash scripts/gpu_ci_bench.sh TAG [workloads...]
TAG=${1:-cur}; shift
mkdir -p gpurun_out
bash scripts/gpu_ci.sh
for w in "${@:-cfg2}"; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary.txt; tail -c 400 gpurun_out/bench_${w}_$TAG.err
  python scripts/show_bench.py gpurun_out/bench_${w}_$TAG.json
done

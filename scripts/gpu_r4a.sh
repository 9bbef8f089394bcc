#!/bin/bash
# round-2 session 3, call a: float2 (H*W % 4 == 2) dynamic-filter paths, dfilt with hoisted X loads, split-K dx of the skinny projections
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dynfilter.py tests/test_gpu_att.py tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4a.log 2>&1
echo "pytest exit=$?"; tail -n 5 gpurun_out/test_r4a.log
for w in cfg2 cfg5 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r4a.json 2> gpurun_out/bench_${w}_r4a.err
  echo "bench $w exit=$?"; python scripts/show_bench.py gpurun_out/bench_${w}_r4a.json | head -12
done

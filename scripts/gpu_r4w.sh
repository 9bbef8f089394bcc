#!/bin/bash
# call w: concurrent branch graphs (the N > 1 structure) on ONE GPU: test + bench with L2S_BENCH_FORCE_SPLIT=1 vs the three-graph structure vs one graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_contract.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8
for w in cfg2 cfg4; do
  L2S_BENCH_FORCE_SPLIT=1 timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4w_b.json 2> gpurun_out/bench_${w}_r4w_b.err
  echo "$w branch graphs exit=$? $(python scripts/show_bench.py gpurun_out/bench_${w}_r4w_b.json | head -1 | cut -d' ' -f2-5)"; grep "bench:" gpurun_out/bench_${w}_r4w_b.err
  L2S_BENCH_FORCE_SPLIT=1 L2S_BENCH_BRANCH_GRAPHS=0 timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4w_3.json 2> gpurun_out/bench_${w}_r4w_3.err
  echo "$w three graphs  exit=$? $(python scripts/show_bench.py gpurun_out/bench_${w}_r4w_3.json | head -1 | cut -d' ' -f2-5)"; grep "bench:" gpurun_out/bench_${w}_r4w_3.err
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4w_1.json 2> gpurun_out/bench_${w}_r4w_1.err
  echo "$w one graph     exit=$? $(python scripts/show_bench.py gpurun_out/bench_${w}_r4w_1.json | head -1 | cut -d' ' -f2-5)"; grep "bench:" gpurun_out/bench_${w}_r4w_1.err
done

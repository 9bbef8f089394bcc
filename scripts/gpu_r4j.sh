#!/bin/bash
# call j: A/B on one box -- 0 / 1 (caption) / 2 (caption + mask head) extra streams
mkdir -p gpurun_out
for w in cfg2 cfg3; do
  for s in 0 1 2 0 1 2; do
    L2S_BENCH_STREAMS=$s timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4j_s$s.json 2> gpurun_out/bench_${w}_r4j_s$s.err
    echo "bench $w streams=$s exit=$?"; python scripts/show_bench.py gpurun_out/bench_${w}_r4j_s$s.json 2>/dev/null | head -1; tail -n 2 gpurun_out/bench_${w}_r4j_s$s.err | cut -c1-200
  done
done

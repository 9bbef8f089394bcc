#!/bin/bash
# call t: stream priorities of the branches (same box, alternating)
mkdir -p gpurun_out
python -c "import torch; print('priority range', torch.cuda.Stream.priority_range())"
for w in cfg2 cfg4; do
for p in "0 0" "-1 0" "-5 0" "-5 -1" "0 -5" "0 0"; do
  set -- $p
  L2S_BENCH_PRIO=$1 L2S_BENCH_PRIO2=$2 timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4t.json 2> gpurun_out/bench_${w}_r4t.err
  echo "$w prio caption=$1 mask=$2 exit=$? $(python scripts/show_bench.py gpurun_out/bench_${w}_r4t.json | head -1 | cut -d' ' -f2-5)"
done
done

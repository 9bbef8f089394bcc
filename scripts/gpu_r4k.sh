#!/bin/bash
# call k (2 GPUs): the split-graph N = 2 path with the branch streams; loss with / without streams at N = 1
mkdir -p gpurun_out
for s in 0 2; do
  L2S_BENCH_STREAMS=$s timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4k_s$s.json 2> gpurun_out/bench_cfg2_r4k_s$s.err
  echo "N=1 streams=$s exit=$?"; grep "loss of the graphed" gpurun_out/bench_cfg2_r4k_s$s.err; python scripts/show_bench.py gpurun_out/bench_cfg2_r4k_s$s.json 2>/dev/null | head -1
done
for w in cfg2 cfg4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --workload $w --steps 30 --warmup 5 --no-res5 --no-components > gpurun_out/bench_${w}_n2_r4k.json 2> gpurun_out/bench_${w}_n2_r4k.err
echo "N=2 $w exit=$?"; grep -i "capture failed\|loss of the graphed\|Error" gpurun_out/bench_${w}_n2_r4k.err | head -5; python scripts/show_bench.py gpurun_out/bench_${w}_n2_r4k.json | head -1
done

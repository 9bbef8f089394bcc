#!/bin/bash
# N = 8 at HEAD (packed flat gradients): cfg2 and cfg4
mkdir -p gpurun_out
for w in cfg2 cfg4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 8 --workload $w --steps 30 --warmup 5 --no-res5 --no-components > gpurun_out/bench_${w}_n8_r3b.json 2> gpurun_out/bench_${w}_n8_r3b.err
  echo "N=8 $w exit=$?"; python scripts/show_bench.py gpurun_out/bench_${w}_n8_r3b.json | grep -E "expr/s"
done

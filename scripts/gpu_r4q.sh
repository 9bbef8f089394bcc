#!/bin/bash
# call q (2 GPUs): N = 2 with the cut at the expression embedding and four gradient groups
mkdir -p gpurun_out
for w in cfg2 cfg4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --workload $w --steps 30 --warmup 5 --no-res5 --no-components --no-cpu-baseline > gpurun_out/bench_${w}_n2_r4q.json 2> gpurun_out/bench_${w}_n2_r4q.err
echo "N=2 $w exit=$?"; grep -i "capture failed\|loss of the graphed\|Error" gpurun_out/bench_${w}_n2_r4q.err | head -5; python scripts/show_bench.py gpurun_out/bench_${w}_n2_r4q.json | head -1
done
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_n1_r4q.json 2> gpurun_out/bench_cfg2_n1_r4q.err
echo "N=1 cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_n1_r4q.json | head -1

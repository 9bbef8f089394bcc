#!/bin/bash
# call y (2 GPUs): N = 2 with one graph per branch and per-branch all-reduces
mkdir -p gpurun_out
for w in cfg2 cfg4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --workload $w --steps 30 --warmup 5 --no-res5 --no-components --no-cpu-baseline > gpurun_out/bench_${w}_n2_r4y.json 2> gpurun_out/bench_${w}_n2_r4y.err
echo "N=2 $w exit=$?"; grep -i "bench:\|Error" gpurun_out/bench_${w}_n2_r4y.err | head -5; python scripts/show_bench.py gpurun_out/bench_${w}_n2_r4y.json | head -1
done
timeout 200 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_n1_r4y.json 2> gpurun_out/bench_cfg2_n1_r4y.err
echo "N=1 cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_n1_r4y.json | head -1

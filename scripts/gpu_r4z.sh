#!/bin/bash
# call z (8 GPUs): N = 8 sanity of the per-branch graph structure
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --workload cfg2 --steps 10 --warmup 3 --no-res5 --no-components --no-cpu-baseline > gpurun_out/bench_cfg2_n8_r4z.json 2> gpurun_out/bench_cfg2_n8_r4z.err
echo "N=8 cfg2 exit=$?"; grep -i "bench:\|Error" gpurun_out/bench_cfg2_n8_r4z.err | head -5; python scripts/show_bench.py gpurun_out/bench_cfg2_n8_r4z.json | head -1

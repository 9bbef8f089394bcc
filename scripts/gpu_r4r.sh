#!/bin/bash
# call r (8 GPUs): N = 8 sanity of the session-3 step (branch streams, cut at the expression embedding, four gradient groups)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 8 --workload cfg2 --steps 30 --warmup 5 --no-res5 --no-components --no-cpu-baseline > gpurun_out/bench_cfg2_n8_r4r.json 2> gpurun_out/bench_cfg2_n8_r4r.err
echo "N=8 cfg2 exit=$?"; grep -i "capture failed\|loss of the graphed\|Error" gpurun_out/bench_cfg2_n8_r4r.err | head -5; python scripts/show_bench.py gpurun_out/bench_cfg2_n8_r4r.json | head -1

#!/bin/bash
# multi-GPU: flat gradients + overlapped all-reduce (N = 2 here; the driver runs 1/2/4/8 at round end)
mkdir -p gpurun_out
N=${1:-2}
for w in cfg2 cfg4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-components > gpurun_out/bench_${w}_n${N}_r2i.json 2> gpurun_out/bench_${w}_n${N}_r2i.err
  echo "bench $w N=$N exit=$?" | tee -a gpurun_out/summary_r2i.txt; grep -v "^$" gpurun_out/bench_${w}_n${N}_r2i.err | tail -n 5
  python scripts/show_bench.py gpurun_out/bench_${w}_n${N}_r2i.json
done
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_cfg2_n1_r2i.json 2> gpurun_out/bench_cfg2_n1_r2i.err
python scripts/show_bench.py gpurun_out/bench_cfg2_n1_r2i.json

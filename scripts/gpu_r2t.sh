#!/bin/bash
# where do the 0.4-0.5 ms of the N > 1 step go?  N = 2: as shipped / without the collectives / NCCL limited to few CTAs
mkdir -p gpurun_out
run() {  # tag env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus 2 --workload cfg2 --steps 30 --warmup 5 --no-res5 --no-components > gpurun_out/bench_cfg2_n2_$tag.json 2> gpurun_out/bench_cfg2_n2_$tag.err
  echo "N=2 $tag exit=$?"; tail -c 300 gpurun_out/bench_cfg2_n2_$tag.err | tail -n 2
  python scripts/show_bench.py gpurun_out/bench_cfg2_n2_$tag.json | grep -E "expr/s"
}
run r2t_shipped L2S_X=0
run r2t_nocomm L2S_BENCH_NOCOMM=1
run r2t_ctas4 NCCL_MAX_CTAS=4
run r2t_ctas16 NCCL_MAX_CTAS=16
timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-res5 --no-components --no-cpu-baseline > gpurun_out/bench_cfg2_n1_r2t.json 2>/dev/null
python scripts/show_bench.py gpurun_out/bench_cfg2_n1_r2t.json | grep -E "expr/s"

#!/bin/bash
# N = 2 with packed flat gradients (one fused copy per group instead of ~100 accumulate kernels)
mkdir -p gpurun_out
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus 2 --workload $wl --steps 30 --warmup 5 --no-res5 --no-components > gpurun_out/bench_${wl}_n2_$tag.json 2> gpurun_out/bench_${wl}_n2_$tag.err
  echo "N=2 $wl $tag exit=$?"; tail -c 300 gpurun_out/bench_${wl}_n2_$tag.err | tail -n 2
  python scripts/show_bench.py gpurun_out/bench_${wl}_n2_$tag.json | grep -E "expr/s"
}
run r2u cfg2 L2S_X=0
run r2u_nocomm cfg2 L2S_BENCH_NOCOMM=1
run r2u cfg4 L2S_X=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --no-components > gpurun_out/bench_cfg2_n2_r2u_res5.json 2> gpurun_out/bench_cfg2_n2_r2u_res5.err
echo "N=2 default (with res5) exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_n2_r2u_res5.json | grep -E "expr/s"

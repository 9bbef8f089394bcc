#!/bin/bash
# RoI max-pool B200 kernels (smem-resident forward, pipelined owner backward): parity + timing; store-policy A/B of the dynamic filter gating
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2q.txt
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q -m gpu -p no:cacheprovider -k "maxpool" > gpurun_out/test_r2q.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2q.txt; tail -n 8 gpurun_out/test_r2q.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg2_r2q.json 2> gpurun_out/bench_cfg2_r2q.err
echo "bench cfg2 exit=$?" | tee -a gpurun_out/summary_r2q.txt; tail -c 300 gpurun_out/bench_cfg2_r2q.err
python scripts/show_bench.py gpurun_out/bench_cfg2_r2q.json | grep -E "expr/s|dynfilter|maxpool"
for v in 1 2; do
  L2S_NVCC_FLAGS="-DL2S_DT_STORE=$v" python -m lang2seg_b200.build --force > gpurun_out/build_store$v.log 2>&1
  echo "build store=$v exit=$?" | tee -a gpurun_out/summary_r2q.txt
  for w in cfg2 cfg4; do
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r2q_store$v.json 2> gpurun_out/bench_${w}_r2q_store$v.err
    python scripts/show_bench.py gpurun_out/bench_${w}_r2q_store$v.json | grep -E "expr/s|dynfilter_fwd"
  done
done

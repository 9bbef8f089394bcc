#!/bin/bash
# owner-computes RoI max-pool backward, dfilt with float masks, bwd ring depth: parity + cfg2 bench (A/B atomic max-pool) + launch list
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2p.txt
timeout 900 python -m pytest tests/test_gpu_roi.py tests/test_gpu_dynfilter.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_r2p.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2p.txt; tail -n 8 gpurun_out/test_r2p.log
for mode in owner atomic; do
  if [ $mode = atomic ]; then export L2S_ROIPOOL_BWD_ATOMIC=1; else unset L2S_ROIPOOL_BWD_ATOMIC; fi
  timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg2_r2p_$mode.json 2> gpurun_out/bench_cfg2_r2p_$mode.err
  echo "bench cfg2 $mode exit=$?" | tee -a gpurun_out/summary_r2p.txt; tail -c 300 gpurun_out/bench_cfg2_r2p_$mode.err
  python scripts/show_bench.py gpurun_out/bench_cfg2_r2p_$mode.json | grep -E "expr/s|dynfilter|maxpool"
done
unset L2S_ROIPOOL_BWD_ATOMIC
timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg4_r2p.json 2> gpurun_out/bench_cfg4_r2p.err
python scripts/show_bench.py gpurun_out/bench_cfg4_r2p.json | grep -E "expr/s|dynfilter"
bash scripts/gpu_launches.sh r2p

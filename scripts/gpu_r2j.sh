#!/bin/bash
# tensor-core dynamic filter forward + proposal kernels: parity, then bench cfg2 / cfg4 with and without the tcgen05 kernel
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2j.txt
timeout 600 python -m pytest tests/test_gpu_dynfilter.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_dynfilter.log 2>&1
echo "test_gpu_dynfilter exit=$?" | tee -a gpurun_out/summary_r2j.txt; tail -n 25 gpurun_out/test_dynfilter.log
timeout 600 python -m pytest tests/test_gpu_proposals.py -q -m gpu -p no:cacheprovider > gpurun_out/test_proposals.log 2>&1
echo "test_gpu_proposals exit=$?" | tee -a gpurun_out/summary_r2j.txt; tail -n 25 gpurun_out/test_proposals.log
for mode in tc ffma; do
  if [ $mode = ffma ]; then export L2S_DYNFILTER_FFMA=1; else unset L2S_DYNFILTER_FFMA; fi
  for w in cfg2 cfg4; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2j_$mode.json 2> gpurun_out/bench_${w}_r2j_$mode.err
    echo "bench $w $mode exit=$?" | tee -a gpurun_out/summary_r2j.txt; tail -c 300 gpurun_out/bench_${w}_r2j_$mode.err
    python scripts/show_bench.py gpurun_out/bench_${w}_r2j_$mode.json | grep -E "expr/s|dynfilter"
  done
done

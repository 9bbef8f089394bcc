#!/bin/bash
# tcgen05 dynamic filter v3 (stacked operands: one MMA per 16 channels; two converter groups): parity, A/B bench, ncu, launch list
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2m.txt
for tpx in 16 32; do
  L2S_DYNFILTER_TPX=$tpx timeout 600 python -m pytest tests/test_gpu_dynfilter.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_dynfilter_tpx$tpx.log 2>&1
  echo "test_gpu_dynfilter tpx=$tpx exit=$?" | tee -a gpurun_out/summary_r2m.txt; tail -n 6 gpurun_out/test_dynfilter_tpx$tpx.log
done
for tpx in 16 32; do
  for w in cfg2 cfg4; do
    L2S_DYNFILTER_TPX=$tpx timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r2m_tpx$tpx.json 2> gpurun_out/bench_${w}_r2m_tpx$tpx.err
    echo "bench $w tpx=$tpx exit=$?" | tee -a gpurun_out/summary_r2m.txt; tail -c 300 gpurun_out/bench_${w}_r2m_tpx$tpx.err
    python scripts/show_bench.py gpurun_out/bench_${w}_r2m_tpx$tpx.json | grep -E "expr/s|dynfilter"
  done
done
for tpx in 16 32; do
  export L2S_DYNFILTER_TPX=$tpx
  bash scripts/gpu_prof_one.sh dyntc3_tpx$tpx "dynfilter_tc_fwd" dyn 1
  python scripts/sass_stalls.py gpurun_out/prof_dyntc3_tpx${tpx}_sass.csv 12 > gpurun_out/prof_dyntc3_tpx${tpx}_stalls.txt 2>&1
  head -24 gpurun_out/prof_dyntc3_tpx${tpx}_stalls.txt
done
unset L2S_DYNFILTER_TPX
bash scripts/gpu_launches.sh r2m

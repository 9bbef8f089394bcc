#!/bin/bash
# TMA-streamed dynamic filter backward + vectorised dfilt: parity, A/B bench; sanitizer runs; launch list without the persistent kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2o.txt
timeout 600 python -m pytest tests/test_gpu_dynfilter.py tests/test_gpu_net.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_r2o.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2o.txt; tail -n 8 gpurun_out/test_r2o.log
for mode in tma ffma; do
  if [ $mode = ffma ]; then export L2S_DYNFILTER_BWD_FFMA=1 L2S_DFILT_SCALAR=1; else unset L2S_DYNFILTER_BWD_FFMA L2S_DFILT_SCALAR; fi
  for w in cfg2 cfg4; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r2o_$mode.json 2> gpurun_out/bench_${w}_r2o_$mode.err
    echo "bench $w $mode exit=$?" | tee -a gpurun_out/summary_r2o.txt; tail -c 300 gpurun_out/bench_${w}_r2o_$mode.err
    python scripts/show_bench.py gpurun_out/bench_${w}_r2o_$mode.json | grep -E "expr/s|dynfilter"
  done
done
unset L2S_DYNFILTER_BWD_FFMA L2S_DFILT_SCALAR
bash scripts/gpu_prof_one.sh dynbwd "dynfilter_bwd_tma|dfilt_vec" dyn 2
python scripts/sass_stalls.py gpurun_out/prof_dynbwd_sass.csv 12 > gpurun_out/prof_dynbwd_stalls.txt 2>&1
head -40 gpurun_out/prof_dynbwd_stalls.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dynfilter.py -x -q -m gpu -p no:cacheprovider -k "vs_oracle and sigmoid" > gpurun_out/sanitize_dynfilter_memcheck.log 2>&1
echo "memcheck dynfilter exit=$?" | tee -a gpurun_out/summary_r2o.txt; tail -n 6 gpurun_out/sanitize_dynfilter_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_crop.py > gpurun_out/sanitize_crop_racecheck.log 2>&1
echo "racecheck crop exit=$?" | tee -a gpurun_out/summary_r2o.txt; tail -n 12 gpurun_out/sanitize_crop_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_crop.py > gpurun_out/sanitize_crop_memcheck.log 2>&1
echo "memcheck crop exit=$?" | tee -a gpurun_out/summary_r2o.txt; tail -n 8 gpurun_out/sanitize_crop_memcheck.log

#!/bin/bash
# round 2, visit C: persistent decode / LSTM kernels (parity + bench), box head, ncu of the generic row-owner crop backward
mkdir -p gpurun_out
for t in att heads net; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit=$?" | tee -a gpurun_out/summary_r2c.txt
  tail -n 12 gpurun_out/test_$t.log
done
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2c.json 2> gpurun_out/bench_${w}_r2c.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2c.txt; tail -c 600 gpurun_out/bench_${w}_r2c.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2c.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_crop_bwd_rowsx|roi_crop_fwd_kernel' -c 4 -f -o gpurun_out/prof_r2c \
    python scripts/prof_ops.py --reps 1 --only cropmax --workload cfg3 > gpurun_out/prof_r2c.log 2>&1
echo "ncu full exit=$?" | tee -a gpurun_out/summary_r2c.txt
ncu -i gpurun_out/prof_r2c.ncu-rep --page raw --csv > gpurun_out/prof_r2c_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2c.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_r2c_sass.csv 2>/dev/null
ls -la gpurun_out/prof_r2c*

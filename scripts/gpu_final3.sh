#!/bin/bash
# closing verification of round 2, session 3: full GPU CI + smoke + the default bench line (N = 1) + the other configs
mkdir -p gpurun_out
rm -f gpurun_out/summary_final3.txt
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_all_final3.log 2>&1
echo "pytest -m gpu exit=$?" | tee -a gpurun_out/summary_final3.txt; tail -n 3 gpurun_out/test_all_final3.log | tee -a gpurun_out/summary_final3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final3.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary_final3.txt; tail -n 2 gpurun_out/smoke_final3.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg2_s3.json 2> gpurun_out/r02_bench_cfg2_s3.err
echo "bench cfg2 (default) exit=$?" | tee -a gpurun_out/summary_final3.txt; tail -c 400 gpurun_out/r02_bench_cfg2_s3.err
python scripts/show_bench.py gpurun_out/r02_bench_cfg2_s3.json
for w in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_s3.json 2> gpurun_out/r02_bench_${w}_s3.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_final3.txt
  python scripts/show_bench.py gpurun_out/r02_bench_${w}_s3.json | head -12
done
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 > gpurun_out/step_kernels_cfg2_s3.txt 2>&1
echo "prof_step exit=$?" | tee -a gpurun_out/summary_final3.txt

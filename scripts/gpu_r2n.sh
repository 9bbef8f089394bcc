#!/bin/bash
# full GPU CI, split-K rasterisation A/B, in-situ kernel breakdown of the cfg2 step
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2n.txt
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_r2n.log 2>&1
echo "pytest -m gpu exit=$?" | tee -a gpurun_out/summary_r2n.txt; tail -n 8 gpurun_out/test_all_r2n.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2n.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary_r2n.txt; tail -n 3 gpurun_out/smoke_r2n.log
for r in 0 1; do
  L2S_GEMM_RASTER=$r timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_cfg2_r2n_raster$r.json 2> gpurun_out/bench_cfg2_r2n_raster$r.err
  echo "bench cfg2 raster=$r exit=$?" | tee -a gpurun_out/summary_r2n.txt; tail -c 300 gpurun_out/bench_cfg2_r2n_raster$r.err
  python scripts/show_bench.py gpurun_out/bench_cfg2_r2n_raster$r.json | grep -E "expr/s|mask_head|dynfilter"
done
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 > gpurun_out/step_kernels_r2n.txt 2>&1
echo "prof_step exit=$?" | tee -a gpurun_out/summary_r2n.txt; head -60 gpurun_out/step_kernels_r2n.txt

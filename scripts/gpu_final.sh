#!/bin/bash
# round-2 closing run: full GPU CI + smoke, the bench line of every BASELINE.json config, the reference (CPU) arm
mkdir -p gpurun_out
rm -f gpurun_out/summary_final.txt
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_all_final.log 2>&1
echo "pytest -m gpu exit=$?" | tee -a gpurun_out/summary_final.txt; tail -n 4 gpurun_out/test_all_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary_final.txt; tail -n 2 gpurun_out/smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err
echo "bench cfg2 (default) exit=$?" | tee -a gpurun_out/summary_final.txt; tail -c 300 gpurun_out/r02_bench_cfg2.err
python scripts/show_bench.py gpurun_out/r02_bench_cfg2.json
for w in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_final.txt; tail -c 300 gpurun_out/r02_bench_$w.err
  python scripts/show_bench.py gpurun_out/r02_bench_$w.json
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cfg2.json 2> gpurun_out/r02_bench_reference_cfg2.err
echo "reference arm exit=$?" | tee -a gpurun_out/summary_final.txt; cut -c1-600 gpurun_out/r02_bench_reference_cfg2.json

#!/bin/bash
# closing verification at HEAD: full GPU CI + smoke + the default bench line (N = 1) and the reference arm
mkdir -p gpurun_out
rm -f gpurun_out/summary_final2.txt
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_all_final2.log 2>&1
echo "pytest -m gpu exit=$?" | tee -a gpurun_out/summary_final2.txt; tail -n 3 gpurun_out/test_all_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final2.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary_final2.txt; tail -n 2 gpurun_out/smoke_final2.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg2_head.json 2> gpurun_out/r02_bench_cfg2_head.err
echo "bench cfg2 (default) exit=$?" | tee -a gpurun_out/summary_final2.txt; tail -c 300 gpurun_out/r02_bench_cfg2_head.err
python scripts/show_bench.py gpurun_out/r02_bench_cfg2_head.json
for w in cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_head.json 2> gpurun_out/r02_bench_${w}_head.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_final2.txt
  python scripts/show_bench.py gpurun_out/r02_bench_${w}_head.json | grep -E "expr/s"
done

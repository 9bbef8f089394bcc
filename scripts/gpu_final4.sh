#!/bin/bash
# closing verification at HEAD (session 3, after the capture-retry / test additions): full GPU CI + smoke + default bench + reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_all_final4.log 2>&1
echo "pytest -m gpu exit=$?"; tail -n 3 gpurun_out/test_all_final4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final4.log 2>&1
echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke_final4.log
timeout 900 python bench.py > gpurun_out/r02_bench_cfg2_final4.json 2> gpurun_out/r02_bench_cfg2_final4.err
echo "bench (no flags) exit=$?"; grep "bench:" gpurun_out/r02_bench_cfg2_final4.err; python scripts/show_bench.py gpurun_out/r02_bench_cfg2_final4.json | head -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final4.json 2> gpurun_out/r02_bench_reference_final4.err
echo "reference arm exit=$?"; cut -c1-400 gpurun_out/r02_bench_reference_final4.json

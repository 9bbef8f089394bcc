#!/bin/bash
# call p: cut at the expression embedding (4 gradient groups): GPU test + the N > 1 structure on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_contract.py tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
for f in 0 1; do
  L2S_BENCH_FORCE_SPLIT=$f timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4p_f$f.json 2> gpurun_out/bench_cfg2_r4p_f$f.err
  echo "force_split=$f exit=$?"; grep "loss of the graphed\|capture failed" gpurun_out/bench_cfg2_r4p_f$f.err; python scripts/show_bench.py gpurun_out/bench_cfg2_r4p_f$f.json 2>/dev/null | head -1
done

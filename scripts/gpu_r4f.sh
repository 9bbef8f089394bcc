#!/bin/bash
# call f: embedding backward v2 (ballots), colsum reverted; per-launch timeline of the small GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_att.py -q -m gpu -p no:cacheprovider -x -k "embedding or colsum or lang_encoder or decode" > gpurun_out/test_r4f.log 2>&1
echo "pytest exit=$?"; tail -n 3 gpurun_out/test_r4f.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4f.json 2> gpurun_out/bench_cfg2_r4f.err
echo "bench cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_r4f.json 2>/dev/null | head -1
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 --trace all > gpurun_out/step_kernels_cfg2_r4f.txt 2>&1
echo "prof_step exit=$?"

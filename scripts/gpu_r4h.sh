#!/bin/bash
# call h: embedding backward v4 (16 warps), one-stage colsum for <= 1024 rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_att.py tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4h.log 2>&1
echo "pytest exit=$?"; tail -n 3 gpurun_out/test_r4h.log
for i in 1 2; do
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4h$i.json 2> gpurun_out/bench_cfg2_r4h$i.err
echo "bench cfg2 exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_r4h$i.json 2>/dev/null | head -1
done
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 --trace embedding,colsum > gpurun_out/step_kernels_cfg2_r4h.txt 2>&1
echo "prof_step exit=$?"; awk '/# timeline/{f=1} f' gpurun_out/step_kernels_cfg2_r4h.txt | cut -c1-150 | head -24

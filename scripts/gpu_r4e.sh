#!/bin/bash
# call e: vectorised bf16 split, embedding backward kernel, filter generator on the tcgen05 GEMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_att.py tests/test_gpu_dynfilter.py tests/test_gpu_mask_head.py tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4e.log 2>&1
echo "pytest exit=$?"; tail -n 4 gpurun_out/test_r4e.log
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r4e.json 2> gpurun_out/bench_${w}_r4e.err
  echo "bench $w exit=$?"; python scripts/show_bench.py gpurun_out/bench_${w}_r4e.json 2>/dev/null | head -1
done
L2S_DENSE_SMALL_FFMA=1 timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4e_ffma.json 2> gpurun_out/bench_cfg2_r4e_ffma.err
echo "bench cfg2 (FFMA filter generator) exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg2_r4e_ffma.json 2>/dev/null | head -1
timeout 600 python scripts/prof_step.py --workload cfg2 --steps 3 > gpurun_out/step_kernels_cfg2_r4e.txt 2>&1
echo "prof_step exit=$?"

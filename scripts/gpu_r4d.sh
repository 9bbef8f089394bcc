#!/bin/bash
# call d: recurrence weight gradients on the tcgen05 GEMM, crop backward reusing the forward's workspace, colsum with more loads in flight
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_roi.py tests/test_gpu_att.py tests/test_gpu_net.py tests/test_gpu_heads.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4d.log 2>&1
echo "pytest exit=$?"; tail -n 4 gpurun_out/test_r4d.log
for w in cfg2 cfg4 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r4d.json 2> gpurun_out/bench_${w}_r4d.err
  echo "bench $w exit=$?"; python scripts/show_bench.py gpurun_out/bench_${w}_r4d.json 2>/dev/null | head -5
done
L2S_WGRAD_FFMA=1 timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg4_r4d_ffma.json 2> gpurun_out/bench_cfg4_r4d_ffma.err
echo "bench cfg4 (FFMA wgrad) exit=$?"; python scripts/show_bench.py gpurun_out/bench_cfg4_r4d_ffma.json 2>/dev/null | head -1

#!/bin/bash
# Runs the GPU parity suites one file per process (a hang or crash in one cannot take the others down).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in roi dynfilter att mask_head targets nms heads net; do
  timeout ${L2S_TEST_TIMEOUT:-420} python -m pytest tests/test_gpu_$t.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/test_$t.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log

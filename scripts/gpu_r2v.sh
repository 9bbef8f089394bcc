#!/bin/bash
# GEMM1 with in-kernel fp32 -> bf16-plane conversion of A (no repack_x pass): parity of the mask head, A/B bench
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2v.txt
timeout 900 python -m pytest tests/test_gpu_mask_head.py tests/test_gpu_net.py tests/test_gpu_heads.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_r2v.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2v.txt; tail -n 12 gpurun_out/test_r2v.log
for mode in fused repack; do
  if [ $mode = repack ]; then export L2S_MASK_REPACK=1; else unset L2S_MASK_REPACK; fi
  for w in cfg2 cfg3; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 > gpurun_out/bench_${w}_r2v_$mode.json 2> gpurun_out/bench_${w}_r2v_$mode.err
    echo "bench $w $mode exit=$?" | tee -a gpurun_out/summary_r2v.txt; tail -c 300 gpurun_out/bench_${w}_r2v_$mode.err
    python scripts/show_bench.py gpurun_out/bench_${w}_r2v_$mode.json | grep -E "expr/s|mask_head|gemm"
  done
done

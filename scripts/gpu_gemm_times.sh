#!/bin/bash
# per-kernel device times of the mask-head / caption GEMMs for every CTA shape (ncu launch list; cold, serialised)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mask_head.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_mask_head_mh.log 2>&1
echo "test_gpu_mask_head exit=$?" | tee gpurun_out/summary_gemm.txt
tail -n 4 gpurun_out/test_mask_head_mh.log
for SH in ${SHAPES:-lib 0 1 2}; do
  if [ "$SH" = lib ]; then unset L2S_GEMM_SHAPE; else export L2S_GEMM_SHAPE=$SH; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_bf16x3|repack|split_kernel' --csv \
      --log-file gpurun_out/gemm_sh$SH.csv python scripts/prof_ops.py --reps 2 --only mask,cap > gpurun_out/gemm_sh$SH.log 2>&1
  echo "ncu shape $SH exit=$?" | tee -a gpurun_out/summary_gemm.txt
  python scripts/summarize_launches.py gpurun_out/gemm_sh$SH.csv | head -20
done

#!/bin/bash
# ncu --set full capture of the hot-path kernels (one GPU).  Usage: bash scripts/gpu_prof.sh [regex] [count] [only]
mkdir -p gpurun_out
RX=${1:-'roi_crop|dynfilter|att_step|gemm_bf16x3|repack'}
CNT=${2:-40}
ONLY=${3:-dyn,crop,cropmax,mask,att}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -f -o gpurun_out/prof \
    python scripts/prof_ops.py --reps 1 --only $ONLY > gpurun_out/prof.log 2>&1
echo "ncu full exit=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/prof.log
ls -la gpurun_out/prof.ncu-rep

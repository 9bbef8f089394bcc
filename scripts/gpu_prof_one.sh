#!/bin/bash
# ncu --set full (+source) of single launches: bash scripts/gpu_prof_one.sh <tag> <kernel regex> <prof_ops --only list> [count]
TAG=$1; RX=$2; ONLY=$3; CNT=${4:-1}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -f -o gpurun_out/prof_$TAG \
    python scripts/prof_ops.py --reps 1 --only $ONLY > gpurun_out/prof_$TAG.log 2>&1
echo "ncu $TAG exit=$?" | tee -a gpurun_out/summary_prof.txt
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${TAG}_sass.csv 2>/dev/null
ls -la gpurun_out/prof_$TAG*

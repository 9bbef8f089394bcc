#!/bin/bash
# launch list of one bench step (ncu, serialised): bash scripts/gpu_launches.sh [tag]
# The persistent cooperative kernels (decode / bi-LSTM: grid barriers) do not survive ncu's serialisation and are left
# out of the list (their in-situ times are in the torch-profiler breakdown, scripts/prof_step.py).
mkdir -p gpurun_out
TAG=${1:-cur}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*persist)' -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-components --no-res5 > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu launches exit=$?" | tee -a gpurun_out/summary.txt
python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.txt
head -50 gpurun_out/launches_${TAG}_summary.txt

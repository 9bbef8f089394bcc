#!/bin/bash
# One GPU visit: parity suites + smoke, bench (cfg2), ncu launch list of the bench command, ncu --set full of the
# hot kernels (raw page exported as CSV on the box).  Usage: bash scripts/gpu_round.sh [tag]
TAG=${1:-cur}
mkdir -p gpurun_out
bash scripts/gpu_ci.sh
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
echo "bench cfg2 exit=$?" | tee -a gpurun_out/summary.txt; tail -c 800 gpurun_out/bench_cfg2_$TAG.err
cat gpurun_out/bench_cfg2_$TAG.json
bash scripts/gpu_launches.sh $TAG
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'roi_crop_(fwd|bwd)|dynfilter_(fwd|bwd)|att_step|gemm_bf16x3|mask_bce_du|repack_x|att_accum' -c 24 -f -o gpurun_out/prof_$TAG \
    python scripts/prof_ops.py --reps 1 --only dyn,crop,maskloss,att > gpurun_out/prof_$TAG.log 2>&1
echo "ncu full exit=$?" | tee -a gpurun_out/summary.txt
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/
# keep the report only if it fits comfortably in the 64 MiB return budget
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; echo "ncu-rep dropped ($SZ bytes)"; fi

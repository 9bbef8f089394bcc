#!/bin/bash
# cfg5 batch sweep (I = E in 8..256 on one GPU) and ncu of the max-pool crop kernels at cfg3 / cfg5 shapes
mkdir -p gpurun_out
for b in 8 16 32 64 128 256; do
  timeout 600 python bench.py --workload cfg5 --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_cfg5_b$b.json 2> gpurun_out/bench_cfg5_b$b.err
  echo "cfg5 batch $b exit=$?"; tail -c 200 gpurun_out/bench_cfg5_b$b.err | tail -n 1
  python scripts/show_bench.py gpurun_out/bench_cfg5_b$b.json | grep -E "expr/s"
done
for w in cfg3 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_crop" -c 6 -f -o gpurun_out/prof_cropmax_$w \
      python scripts/prof_ops.py --reps 1 --only cropmax --workload $w > gpurun_out/prof_cropmax_$w.log 2>&1
  echo "ncu cropmax $w exit=$?"
  ncu -i gpurun_out/prof_cropmax_$w.ncu-rep --page raw --csv > gpurun_out/prof_cropmax_${w}_raw.csv 2>/dev/null
done
python - <<'PY'
import csv
for w in ("cfg3","cfg5"):
    rows=list(csv.reader(open('gpurun_out/prof_cropmax_%s_raw.csv'%w)))
    hdr=rows[0]
    for r in rows[2:]:
        g=lambda k: r[hdr.index(k)] if k in hdr else ""
        print(w, g("Kernel Name")[:70], "dur", g("gpu__time_duration.sum"), "rd", g("dram__bytes_read.sum"), "wr", g("dram__bytes_write.sum"), "smemwf", g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), "issue", g("smsp__issue_active.avg.pct_of_peak_sustained_active"))
PY

#!/bin/bash
# weight gradients of the skinny linears / recurrences on the library's own FFMA GEMM (no cuBLAS on the step): parity + bench;
# ncu of the mask-head GEMMs (DRAM bytes of the split-major weight-gradient GEMM)
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2x.txt
timeout 900 python -m pytest tests/test_gpu_att.py tests/test_gpu_dynfilter.py tests/test_gpu_net.py tests/test_gpu_heads.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_r2x.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2x.txt; tail -n 8 gpurun_out/test_r2x.log
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_${w}_r2x.json 2> gpurun_out/bench_${w}_r2x.err
  echo "bench $w exit=$?" | tee -a gpurun_out/summary_r2x.txt; tail -c 300 gpurun_out/bench_${w}_r2x.err
  python scripts/show_bench.py gpurun_out/bench_${w}_r2x.json | grep -E "expr/s"
done
bash scripts/gpu_prof_one.sh gemms_r2x "gemm_bf16x3" maskloss 4
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_gemms_r2x_raw.csv')))
hdr=rows[0]
for r in rows[2:]:
    g=lambda k: r[hdr.index(k)]
    print(g("Kernel Name")[-60:], "dur", g("gpu__time_duration.sum"), "rd", g("dram__bytes_read.sum"), "wr", g("dram__bytes_write.sum"), "tensor%", g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "L2hit", g("lts__t_sector_hit_rate.pct"))
PY

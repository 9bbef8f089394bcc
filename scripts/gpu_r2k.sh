#!/bin/bash
# proposals + bf16 variants + res5-chained step: parity; ncu of the tcgen05 dynamic filter; default bench line with the chained number
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2k.txt
timeout 900 python -m pytest tests/test_gpu_proposals.py tests/test_gpu_net.py "tests/test_gpu_mask_head.py" -q -m gpu -p no:cacheprovider -k "proposal or chained or hot_path or bf16 or single_pass" > gpurun_out/test_r2k.log 2>&1
echo "tests exit=$?" | tee -a gpurun_out/summary_r2k.txt; tail -n 30 gpurun_out/test_r2k.log
bash scripts/gpu_prof_one.sh dyntc_cfg2 "dynfilter_tc_fwd" dyn 1
python scripts/sass_stalls.py gpurun_out/prof_dyntc_cfg2_sass.csv 25 > gpurun_out/prof_dyntc_cfg2_stalls.txt 2>&1
head -50 gpurun_out/prof_dyntc_cfg2_stalls.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2_r2k.json 2> gpurun_out/bench_cfg2_r2k.err
echo "bench cfg2 default exit=$?" | tee -a gpurun_out/summary_r2k.txt; tail -c 600 gpurun_out/bench_cfg2_r2k.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_r2k.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("with_res5", json.dumps(d.get("with_res5")))
print("cpu", json.dumps(d.get("cpu_baseline")))
PY

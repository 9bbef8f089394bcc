#!/bin/bash
# multi-GPU validation of the round-2 step: N = 8 (default run incl. the res5-chained number), N = 4 and N = 8 on cfg4
mkdir -p gpurun_out
rm -f gpurun_out/summary_r2s.txt
run() {  # n workload extra tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $1 --workload $2 --steps 20 --warmup 3 $3 > gpurun_out/bench_$2_n$1_$4.json 2> gpurun_out/bench_$2_n$1_$4.err
  echo "bench $2 N=$1 $4 exit=$?" | tee -a gpurun_out/summary_r2s.txt; tail -c 400 gpurun_out/bench_$2_n$1_$4.err
  python scripts/show_bench.py gpurun_out/bench_$2_n$1_$4.json | grep -E "expr/s"
}
run 8 cfg2 "" r2s
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_n8_r2s.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "with_res5", json.dumps(d.get("with_res5"))[:400])
PY
run 4 cfg2 "--no-res5 --no-components" r2s
run 8 cfg4 "--no-res5 --no-components" r2s
run 8 cfg3 "--no-res5 --no-components" r2s

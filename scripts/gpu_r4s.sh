#!/bin/bash
# call s: capture retry path (injected failure -> one stream), then the normal path
mkdir -p gpurun_out
L2S_BENCH_TEST_CAPTURE_FAIL=1 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4s_retry.json 2> gpurun_out/bench_cfg2_r4s_retry.err
echo "retry path exit=$?"; grep "bench:" gpurun_out/bench_cfg2_r4s_retry.err; python scripts/show_bench.py gpurun_out/bench_cfg2_r4s_retry.json | head -1
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4s.json 2> gpurun_out/bench_cfg2_r4s.err
echo "normal path exit=$?"; grep "bench:" gpurun_out/bench_cfg2_r4s.err; python scripts/show_bench.py gpurun_out/bench_cfg2_r4s.json | head -1
python -c "
import json
for f in ('gpurun_out/bench_cfg2_r4s_retry.json','gpurun_out/bench_cfg2_r4s.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['config']['streams'], d['config']['launch'])
"

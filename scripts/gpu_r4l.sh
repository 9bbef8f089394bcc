#!/bin/bash
# call l: the N > 1 graph structure (three graphs, backward cut at the generated filters) exercised on ONE GPU; loss must match the one-graph step
mkdir -p gpurun_out
for f in 0 1; do
  L2S_BENCH_FORCE_SPLIT=$f timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg2_r4l_f$f.json 2> gpurun_out/bench_cfg2_r4l_f$f.err
  echo "force_split=$f exit=$?"; grep "loss of the graphed\|capture failed" gpurun_out/bench_cfg2_r4l_f$f.err; python scripts/show_bench.py gpurun_out/bench_cfg2_r4l_f$f.json 2>/dev/null | head -1
done
L2S_BENCH_FORCE_SPLIT=1 timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline --no-res5 --no-components > gpurun_out/bench_cfg4_r4l_f1.json 2> gpurun_out/bench_cfg4_r4l_f1.err
echo "cfg4 force_split exit=$?"; grep "loss of the graphed\|capture failed" gpurun_out/bench_cfg4_r4l_f1.err; python scripts/show_bench.py gpurun_out/bench_cfg4_r4l_f1.json 2>/dev/null | head -1
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_r4l.log 2>&1
echo "pytest exit=$?"; tail -n 2 gpurun_out/test_r4l.log

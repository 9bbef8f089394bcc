#!/bin/bash
# A/B of programmatic dependent launch on the decode / LSTM kernel chains
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_att.py tests/test_gpu_net.py tests/test_gpu_dynfilter.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for P in 0 1; do
  L2S_PDL=$P timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_pdl$P.json 2> gpurun_out/bench_pdl$P.err
  echo "bench pdl=$P exit=$?"; tail -c 300 gpurun_out/bench_pdl$P.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_pdl$P.json')); print('pdl=$P ms/step %.3f value %.1f launch=%s' % (d['ms_per_step'], d['value'], d['config']['launch']))"
done

#!/bin/bash
# call c: fresh in-situ per-kernel breakdown of the cfg2 / cfg4 / cfg5 steps at HEAD
mkdir -p gpurun_out
for w in cfg2 cfg4 cfg5; do
  timeout 600 python scripts/prof_step.py --workload $w --steps 3 > gpurun_out/step_kernels_${w}_r4c.txt 2>&1
  echo "prof_step $w exit=$?"; head -4 gpurun_out/step_kernels_${w}_r4c.txt | cut -c1-160
done

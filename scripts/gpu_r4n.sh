#!/bin/bash
# compute-sanitizer memcheck on the kernels added / changed in session 3
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
( timeout 400 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_att.py -q -m gpu -p no:cacheprovider -x -k "embedding or colsum" ; echo "memcheck embedding/colsum exit=$?" ) > gpurun_out/sanitize_r4n_a.log 2>&1
tail -n 4 gpurun_out/sanitize_r4n_a.log
( timeout 500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dynfilter.py -q -m gpu -p no:cacheprovider -x -k "cfg8 or cfg9 or cfg10 or filter_generator" ; echo "memcheck dynfilter quads / filter generator exit=$?" ) > gpurun_out/sanitize_r4n_b.log 2>&1
tail -n 4 gpurun_out/sanitize_r4n_b.log
( timeout 400 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_roi.py -q -m gpu -p no:cacheprovider -x -k "workspace_reuse" ; echo "memcheck crop workspace reuse exit=$?" ) > gpurun_out/sanitize_r4n_c.log 2>&1
tail -n 4 gpurun_out/sanitize_r4n_c.log
grep -h "ERROR SUMMARY" gpurun_out/sanitize_r4n_*.log

#!/bin/bash
# ncu of the fused GEMM1 (in-kernel A conversion)
mkdir -p gpurun_out
bash scripts/gpu_prof_one.sh gemm1f "gemm_bf16x3" mask 1
python scripts/sass_stalls.py gpurun_out/prof_gemm1f_sass.csv 30 > gpurun_out/prof_gemm1f_stalls.txt 2>&1
head -60 gpurun_out/prof_gemm1f_stalls.txt

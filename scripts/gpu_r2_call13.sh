#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 6" "2 8" "2 10" "3 6" "4 5" "1 16"; do
  set -- $cfg
  echo "=== UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2"
  UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2 timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6
done > gpurun_out/r2_stage_sweep2.txt 2>&1
cat gpurun_out/r2_stage_sweep2.txt

#!/bin/bash
mkdir -p gpurun_out
{
for st in 0 10 12; do
  echo "=== UMB_GEMM_STAGES=$st"
  UMB_GEMM_STAGES=$st timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---" | head -12
done
} > gpurun_out/r2_call38.txt 2>&1
cat gpurun_out/r2_call38.txt

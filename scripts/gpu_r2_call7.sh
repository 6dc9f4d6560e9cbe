#!/bin/bash
mkdir -p gpurun_out
timeout 200 scripts/ubench/tc_rates 2>&1 | sed -n '/against concurrent/,$p' > gpurun_out/r2_tc_contention.txt
cat gpurun_out/r2_tc_contention.txt
for v in "" mma1 mma4 nodq nost nodqnost nodqmma1; do
  if [ -n "$v" ]; then export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so; else unset UMB_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6
done > gpurun_out/r2_exp_variants.txt 2>&1
cat gpurun_out/r2_exp_variants.txt

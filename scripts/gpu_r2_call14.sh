#!/bin/bash
mkdir -p gpurun_out
for v in dq2 dq4; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6
done > gpurun_out/r2_dqgroups.txt 2>&1
cat gpurun_out/r2_dqgroups.txt

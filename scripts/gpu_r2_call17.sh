#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/ubench/tma_issue > gpurun_out/r2_tma_issue.txt 2>&1
cat gpurun_out/r2_tma_issue.txt
for v in serissue parissue; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|gemm_f16|rror" | head -7
  timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
done > gpurun_out/r2_parissue.txt 2>&1
cat gpurun_out/r2_parissue.txt

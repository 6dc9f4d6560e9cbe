#!/bin/bash
mkdir -p gpurun_out
{
echo "=== product: skinny v1b (64 regs, M+1 rows)"
timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
ROWS=2 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
UMB_SKINNY=0 ROWS=2 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
for v in w2 w2x w2x4 x4; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -12
done
} > gpurun_out/r2_call23.txt 2>&1
cat gpurun_out/r2_call23.txt

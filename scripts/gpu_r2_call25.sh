#!/bin/bash
mkdir -p gpurun_out
{
for v in x3d x3dx x3dxw2 x3dxw2q4; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -12
done
} > gpurun_out/r2_call25.txt 2>&1
cat gpurun_out/r2_call25.txt

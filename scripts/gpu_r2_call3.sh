#!/bin/bash
mkdir -p gpurun_out
for v in xskip4 xskip1000; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant '$v'"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|Error|error" | head -20
done > gpurun_out/r2_xskip_b.txt 2>&1
cat gpurun_out/r2_xskip_b.txt
unset UMB_LIB_PATH
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_engine.py -m gpu -q > gpurun_out/r2_pytest3.txt 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/r2_pytest3.txt | tail -20

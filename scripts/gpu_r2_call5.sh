#!/bin/bash
mkdir -p gpurun_out
timeout 300 scripts/ubench/tma_rates > gpurun_out/r2_tma_rates.txt 2>&1
echo "tma rc=$?"; cat gpurun_out/r2_tma_rates.txt
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.rs8576.so
echo "=== variant rs8576" > gpurun_out/r2_rs8576.txt
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|Error|error" | head -8 >> gpurun_out/r2_rs8576.txt
cat gpurun_out/r2_rs8576.txt
unset UMB_LIB_PATH
timeout 900 python -m pytest tests/test_gpu_sampling.py -m gpu -q > gpurun_out/r2_pytest5.txt 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/r2_pytest5.txt | tail -20

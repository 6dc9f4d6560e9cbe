#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2_pytest8.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest8.txt
for cfg in "2 6 12" "2 6 6" "2 6 16" "1 10 10" "2 5 16" "4 3 12"; do
  set -- $cfg
  echo "=== UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2 UMB_GEMM_XSLABS=$3"
  UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2 UMB_GEMM_XSLABS=$3 timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6
done > gpurun_out/r2_xslab_sweep.txt 2>&1
cat gpurun_out/r2_xslab_sweep.txt
timeout 300 python scripts/gemm_chain_trace.py > gpurun_out/r2_chain_trace.txt 2>&1
cat gpurun_out/r2_chain_trace.txt

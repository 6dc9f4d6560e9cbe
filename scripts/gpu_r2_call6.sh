#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2_pytest6.txt 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest6.txt
for cfg in "2 6" "1 10" "4 3" "3 4" "2 8" "4 4"; do
  set -- $cfg
  echo "=== UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2"
  UMB_GEMM_WREQ=$1 UMB_GEMM_STAGES=$2 timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -11
done > gpurun_out/r2_wreq_sweep.txt 2>&1
cat gpurun_out/r2_wreq_sweep.txt
timeout 900 python baseline/ref_gpu.py --steps 10 --warmup 3 > gpurun_out/r2_refgpu.txt 2> gpurun_out/r2_refgpu.err
echo "refgpu rc=$?"; tail -3 gpurun_out/r2_refgpu.txt; tail -15 gpurun_out/r2_refgpu.err

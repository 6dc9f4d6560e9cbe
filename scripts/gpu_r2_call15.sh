#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2_pytest15.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest15.txt
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6 > gpurun_out/r2_xdq.txt
cat gpurun_out/r2_xdq.txt
timeout 200 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace_drain4.txt 2>&1
tail -2 gpurun_out/r2_unit_trace_drain4.txt | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace.txt 2>&1
echo "trace rc=$?"; cat gpurun_out/r2_unit_trace.txt
M=64 timeout 300 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace_m64.txt 2>&1
grep -E "steady|==" gpurun_out/r2_unit_trace_m64.txt
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_mirrors.py tests/test_gpu_offload.py -m gpu -q > gpurun_out/r2_pytest4.txt 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/r2_pytest4.txt | tail -20

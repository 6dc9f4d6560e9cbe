#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/r2_pytest10a.txt 2>&1
echo "pytest gemm rc=$?"; tail -3 gpurun_out/r2_pytest10a.txt
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_parity70b.py tests/test_gpu_engine.py tests/test_gpu_offload.py tests/test_gpu_sampling.py tests/test_gpu_dynamic.py -m gpu -x -q -s > gpurun_out/r2_pytest10b.txt 2>&1
echo "pytest rest rc=$?"; grep -E "parity70b|passed|failed|rror|assert" gpurun_out/r2_pytest10b.txt | tail -12
timeout 200 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace_drain.txt 2>&1
grep -B2 -A60 "gate|up M=13" gpurun_out/r2_unit_trace_drain.txt | head -75

#!/bin/bash
mkdir -p gpurun_out
{
echo "=== final guard form: stress"
M=128 REPS=3000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -5
M=97 REPS=2000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -5
M=65 REPS=1000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -18
timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
} > gpurun_out/r2_call32.txt 2>&1
cat gpurun_out/r2_call32.txt

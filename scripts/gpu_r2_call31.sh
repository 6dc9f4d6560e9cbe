#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest31.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest31.txt
{
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|gemm_f16|tree_attn|rror" | head -40
timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
WIDTH=16 DEPTH=24 BEAMS=24 timeout 600 python scripts/bench_dynamic.py 2>&1 | tail -5; WIDTH=8 DEPTH=8 BEAMS=8 timeout 600 python scripts/bench_dynamic.py 2>&1 | tail -5
} > gpurun_out/r2_call31.txt 2>&1
cat gpurun_out/r2_call31.txt

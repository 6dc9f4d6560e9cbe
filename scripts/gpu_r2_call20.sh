#!/bin/bash
mkdir -p gpurun_out
{
echo "=== product (skinny fp16 linear on)"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
echo "=== product, UMB_SKINNY=0"
UMB_SKINNY=0 timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.cta2.so
echo "=== variant cta2"
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|gemm_f16|rror" | head -24
timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
} > gpurun_out/r2_call20.txt 2>&1
cat gpurun_out/r2_call20.txt

#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python scripts/bench_draft_linears.py 2>&1 | tail -16
UMB_SKINNY=0 timeout 300 python scripts/bench_draft_linears.py 2>&1 | tail -16
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
} > gpurun_out/r2_call21.txt 2>&1
cat gpurun_out/r2_call21.txt

#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_attn.py -m gpu -q -x 2>&1 | tail -12
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "tree_attn"
UMB_ATTN_TC=0 timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "tree_attn"
} > gpurun_out/r2_call40.txt 2>&1
cat gpurun_out/r2_call40.txt

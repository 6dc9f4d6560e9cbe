#!/bin/bash
mkdir -p gpurun_out
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tc.so
{
timeout 200 python -m pytest tests/test_gpu_attn.py -m gpu -q -x 2>&1 | tail -3
echo "=== tcgen05 kernel, 64-row tiles, 5 stages"; timeout 150 python scripts/bench_attn_layout.py 2>&1 | grep "NHD"
} > gpurun_out/r2_call47.txt 2>&1
cat gpurun_out/r2_call47.txt

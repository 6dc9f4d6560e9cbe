#!/bin/bash
mkdir -p gpurun_out
{
echo "=== mma.sync kernels"; timeout 150 python scripts/bench_attn_layout.py 2>&1 | grep tree_attn
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tc.so
echo "=== tcgen05 kernel"; timeout 150 python scripts/bench_attn_layout.py 2>&1 | grep tree_attn
} > gpurun_out/r2_attn_layout.txt 2>&1
cat gpurun_out/r2_attn_layout.txt

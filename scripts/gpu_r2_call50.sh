#!/bin/bash
mkdir -p gpurun_out
{
echo "=== default threshold (tcgen05 kernel from 8192 rows)"; timeout 150 python scripts/bench_attn_ctx.py 2>&1 | grep kv=
echo "=== UMB_ATTN_TC_MIN_KV=256"; UMB_ATTN_TC_MIN_KV=256 timeout 150 python scripts/bench_attn_ctx.py 2>&1 | grep kv=
} > gpurun_out/r2_attn_ctx.txt 2>&1
cat gpurun_out/r2_attn_ctx.txt

#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_attn.py tests/test_gpu_forward.py tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
UMB_ATTN_WARPS=8 timeout 900 python -m pytest tests/test_gpu_attn.py -m gpu -q -x 2>&1 | tail -2
for w in 4 8 0; do
  echo "=== UMB_ATTN_WARPS=$w (0 = heuristic)"
  UMB_ATTN_WARPS=$w timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "tree_attn"
  UMB_ATTN_WARPS=$w timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
  UMB_ATTN_WARPS=$w timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -3
done
} > gpurun_out/r2_call36.txt 2>&1
cat gpurun_out/r2_call36.txt

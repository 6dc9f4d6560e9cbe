#!/bin/bash
for i in 1 2; do
echo "=== main"; timeout 120 python scripts/bench_gemm_dbg.py 2>&1 | grep -E "awq_gemm|layer sum"
echo "=== peek"; UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.peek.so timeout 120 python scripts/bench_gemm_dbg.py 2>&1 | grep -E "awq_gemm|layer sum"
done
UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.peek.so python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -2

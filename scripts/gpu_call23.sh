#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_mirrors.py > gpurun_out/tests23.log 2>&1
grep -E "^==|passed|failed" gpurun_out/tests23.log
timeout 200 python scripts/gemm_trace.py 2>&1 | grep -E "^==|acc1|epi1_done|red_" | head -24
timeout 120 python scripts/bench_gemm_dbg.py 2>&1 | grep -E "awq_gemm|layer sum"
timeout 300 python scripts/iter_breakdown.py 2>&1 | grep -E "graph=|graph replay"

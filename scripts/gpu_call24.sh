#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/tests24.log 2>&1; tail -3 gpurun_out/tests24.log
timeout 200 python scripts/gemm_trace.py > gpurun_out/gemm_trace24.log 2>&1; grep -E "^==|acc1|epi1_done|red_" gpurun_out/gemm_trace24.log | head -24
timeout 120 python scripts/bench_gemm_dbg.py 2>&1 | grep -E "awq_gemm|layer sum"
timeout 300 python scripts/iter_breakdown.py 2>&1 | grep -E "graph=|graph replay"

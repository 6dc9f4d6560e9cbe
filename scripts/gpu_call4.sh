#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_dynamic.py > gpurun_out/tests4.log 2>&1
grep -E "^==|passed|failed" gpurun_out/tests4.log
: > gpurun_out/sweep4.log
timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep4.log 2>&1
M=31 timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep4.log 2>&1
UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.dq4.so timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep4.log 2>&1
F16=1 M=3 timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep4.log 2>&1
grep -E "config|layer sum|awq_gemm|gemm_f16" gpurun_out/sweep4.log
timeout 120 python scripts/gemm_trace.py > gpurun_out/gemm_trace4.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 1500 gpurun_out/bench4.json

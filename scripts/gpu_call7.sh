#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash scripts/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py > gpurun_out/tests7.log 2>&1
grep -E "^==|passed|failed" gpurun_out/tests7.log
: > gpurun_out/sweep7.log
echo "=== main (early reduce)" >> gpurun_out/sweep7.log
timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep7.log 2>&1
echo "=== noearly" >> gpurun_out/sweep7.log
UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.noearly.so timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep7.log 2>&1
grep -E "===|layer sum|awq_gemm" gpurun_out/sweep7.log
timeout 400 python scripts/iter_breakdown.py > gpurun_out/iter_breakdown.log 2>&1; tail -12 gpurun_out/iter_breakdown.log
UMB_DEBUG=1 UMB_PDL=0 LAYERS=80 ITERS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv python scripts/ncu_iteration.py > gpurun_out/ncu_iter7.log 2>&1; grep -E "argument check|done|Error" gpurun_out/ncu_iter7.log | head -5

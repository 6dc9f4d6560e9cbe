#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_dynamic.py tests/test_gpu_offload.py > gpurun_out/tests12.log 2>&1
grep -E "^==|passed|failed|Error" gpurun_out/tests12.log | head -20
timeout 300 python scripts/iter_breakdown.py > gpurun_out/iter_fused.log 2>&1; tail -8 gpurun_out/iter_fused.log
UMB_FUSE_ROPE=0 timeout 300 python scripts/iter_breakdown.py > gpurun_out/iter_unfused.log 2>&1; tail -8 gpurun_out/iter_unfused.log

#!/bin/bash
mkdir -p gpurun_out
M=128 REPS=300 timeout 200 python scripts/repro_m128.py 2>&1 | grep -E "ok|FAILED" | head -4
M=97 REPS=200 timeout 200 python scripts/repro_m128.py 2>&1 | grep -E "ok|FAILED" | head -4
for cfg in "GRAPH=0 DECODE=0" "GRAPH=1 DECODE=3" "GRAPH=0 DECODE=0 UMB_PDL=0"; do
  env $cfg REPS=80 timeout 400 python scripts/stress_prefill.py 2>&1 | grep -E "FAILED|stress done" | tail -2
done
bash scripts/gpu_tests.sh tests/test_gpu_tree.py tests/test_gpu_glue.py tests/test_gpu_attn.py tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_dynamic.py tests/test_gpu_offload.py > gpurun_out/tests21.log 2>&1
grep -E "^==|passed|failed" gpurun_out/tests21.log

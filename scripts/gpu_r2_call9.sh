#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_parity70b.py tests/test_gpu_engine.py -m gpu -x -q -s > gpurun_out/r2_pytest9.txt 2>&1
echo "pytest rc=$?"; grep -E "parity70b|passed|failed|rror|assert" gpurun_out/r2_pytest9.txt | tail -12
for v in "" nodrain; do
  if [ -n "$v" ]; then export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so; else unset UMB_LIB_PATH; fi
  echo "=== variant '${v:-product (drain)}'"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror"
done > gpurun_out/r2_drain_ab.txt 2>&1
cat gpurun_out/r2_drain_ab.txt
unset UMB_LIB_PATH
timeout 300 python scripts/gemm_chain_trace.py > gpurun_out/r2_chain_trace_drain.txt 2>&1
grep -A17 "awq o M" gpurun_out/r2_chain_trace_drain.txt

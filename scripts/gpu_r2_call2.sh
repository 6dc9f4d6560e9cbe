#!/bin/bash
# round 2, call 2: is the AWQ GEMM bound by the L2->SM traffic of the x slabs?  (timing-only x-skip variants)
mkdir -p gpurun_out
for v in "" xskip4 xskip1000; do
  if [ -n "$v" ]; then export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so; else unset UMB_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---"
done > gpurun_out/r2_xskip.txt 2>&1
cat gpurun_out/r2_xskip.txt
unset UMB_LIB_PATH
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_engine.py tests/test_gpu_tree.py tests/test_gpu_dynamic.py -m gpu -q -s > gpurun_out/r2_pytest2.txt 2>&1
echo "pytest rc=$?"; grep -E "TV distance|passed|failed|Error|assert" gpurun_out/r2_pytest2.txt | tail -30

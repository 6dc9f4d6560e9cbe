#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -m gpu -x -q -s > gpurun_out/r2_pytest11.txt 2>&1
echo "pytest rc=$?"; grep -E "forward parity|passed|failed|rror" gpurun_out/r2_pytest11.txt | tail
timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -6 > gpurun_out/r2_pairdrain.txt
cat gpurun_out/r2_pairdrain.txt
timeout 200 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace_drain2.txt 2>&1
grep -B2 -A45 "gate|up M=13" gpurun_out/r2_unit_trace_drain2.txt | head -52

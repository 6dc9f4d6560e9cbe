#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/ubench/hmma_rate > gpurun_out/r2_hmma_rate.txt 2>&1; cat gpurun_out/r2_hmma_rate.txt
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.cta2.so
{
echo "--- plain"; timeout 120 python scripts/one_gemm.py 13 512 256 2>&1 | tail -3
echo "--- PDL off"; UMB_PDL=0 timeout 120 python scripts/one_gemm.py 13 512 256 2>&1 | tail -3
echo "--- sanitizer"; timeout 600 compute-sanitizer --tool memcheck python scripts/one_gemm.py 13 512 256 2>&1 | grep -v "^$" | head -40
} > gpurun_out/r2_cta2_debug.txt 2>&1
cat gpurun_out/r2_cta2_debug.txt

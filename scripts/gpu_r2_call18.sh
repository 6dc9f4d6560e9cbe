#!/bin/bash
mkdir -p gpurun_out
for v in cta2; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|gemm_f16|rror" | head -24
  timeout 300 python scripts/iter_breakdown.py 2>&1 | tail -8
done > gpurun_out/r2_cta2.txt 2>&1
unset UMB_LIB_PATH
cat gpurun_out/r2_cta2.txt
timeout 900 python bench.py --workload static8b --steps 64 --warmup 5 > gpurun_out/r2_bench_static8b.json 2> gpurun_out/r2_bench_static8b.err
echo "static8b rc=$?"; cat gpurun_out/r2_bench_static8b.json; tail -3 gpurun_out/r2_bench_static8b.err
timeout 1200 python bench.py --workload stochastic32k --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k.json 2> gpurun_out/r2_bench_stochastic32k.err
echo "stochastic32k rc=$?"; cat gpurun_out/r2_bench_stochastic32k.json; tail -3 gpurun_out/r2_bench_stochastic32k.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest16.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest16.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.txt 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.txt
timeout 1500 python bench.py --steps 64 --warmup 5 > gpurun_out/r2_bench_static70b.json 2> gpurun_out/r2_bench_static70b.err
echo "bench rc=$?"; cat gpurun_out/r2_bench_static70b.json; tail -5 gpurun_out/r2_bench_static70b.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest26.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest26.txt
timeout 900 python bench.py --steps 64 --warmup 5 --no-reference-gpu > gpurun_out/r2_bench_static70b_call26.json 2> gpurun_out/r2_bench26.err
echo "bench rc=$?"; cat gpurun_out/r2_bench_static70b_call26.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['verify_pass'])"
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.m128.so
{
echo "=== m128 plain"
M=128 REPS=600 timeout 600 python scripts/repro_m128.py 2>&1 | tail -6
M=97 REPS=600 timeout 600 python scripts/repro_m128.py 2>&1 | tail -6
echo "=== m128 under compute-sanitizer memcheck"
M=128 REPS=150 timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python scripts/repro_m128.py 2>&1 | grep -v "^$" | head -60
echo "=== m128 under compute-sanitizer synccheck"
M=128 REPS=60 timeout 900 compute-sanitizer --tool synccheck --print-limit 3 python scripts/repro_m128.py 2>&1 | grep -v "^$" | head -40
} > gpurun_out/r2_m128_debug.txt 2>&1
cat gpurun_out/r2_m128_debug.txt

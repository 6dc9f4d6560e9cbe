#!/bin/bash
mkdir -p gpurun_out
{
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.m128.so
echo "=== m128 (fixed) plain stress"
M=128 REPS=3000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -5
M=97 REPS=3000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_dynamic.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 32 --warmup 3 --no-reference-gpu --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('m128 bench', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])"
unset UMB_LIB_PATH
timeout 600 python bench.py --steps 32 --warmup 3 --no-reference-gpu --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('product bench', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])"
} > gpurun_out/r2_call30.txt 2>&1
cat gpurun_out/r2_call30.txt

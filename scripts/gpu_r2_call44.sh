#!/bin/bash
mkdir -p gpurun_out
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tc.so
{
timeout 240 python -m pytest tests/test_gpu_attn.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python scripts/bench_kernels.py 2>&1 | grep -E "tree_attn"
timeout 400 python bench.py --workload stochastic32k_hbm --steps 32 --warmup 3 --no-reference-gpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stochastic32k_hbm tc', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'])"
} > gpurun_out/r2_call44.txt 2>&1
cat gpurun_out/r2_call44.txt

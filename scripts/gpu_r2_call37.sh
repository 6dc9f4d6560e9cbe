#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --workload static70b_b200tree --steps 64 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2_bench_static70b_b200tree.json 2>/dev/null
echo "b200tree rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_static70b_b200tree.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['config']['mean_accept_len'], d['config']['expected_accept_len'], d['verify_pass'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -2
timeout 400 python bench.py --workload stochastic32k_hbm --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k_hbm.json 2>/dev/null
echo "32k rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_stochastic32k_hbm.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_final.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_final.txt
timeout 500 python bench.py --workload stochastic32k_hbm --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k_hbm.json 2>/dev/null
echo "32k rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_stochastic32k_hbm.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'])"
timeout 150 python scripts/bench_attn_layout.py 2>&1 | grep NHD

#!/bin/bash
# round-2 evidence: bench lines of every workload (reference GPU arm + CPU baseline + our paged target on the default one),
# ncu --set full of the hot kernels, launch list of one iteration
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 64 --warmup 5 --with-offload > gpurun_out/r2_bench_static70b_final.json 2> gpurun_out/r2_bench_final.err
echo "static70b rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_static70b_final.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['verify_pass'], {k:v for k,v in d.get('reference_gpu',{}).items() if k in ('value','ms_per_step','ratio_ours_over_reference_gpu','unavailable')}, d.get('offload_num_cache_layers_0'), d.get('cpu_baseline',{}).get('value'))"; tail -3 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; tail -c 600 gpurun_out/r2_bench_reference_arm.json
timeout 900 python bench.py --workload static8b --steps 64 --warmup 5 > gpurun_out/r2_bench_static8b.json 2>/dev/null
echo "static8b rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_static8b.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'], d.get('reference_gpu',{}).get('value'))"
timeout 900 python bench.py --workload stochastic32k_hbm --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k_hbm.json 2>/dev/null
echo "stochastic32k_hbm rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_stochastic32k_hbm.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'])"
timeout 1200 python bench.py --workload stochastic32k --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k.json 2>/dev/null
echo "stochastic32k rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_stochastic32k.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['verify_pass'], d.get('kv_paging'))"
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/r2_kernels_full -f python scripts/ncu_gemm.py > gpurun_out/r2_ncu_kernels.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2_kernels_full.ncu-rep gpurun_out/r2_ncu_kernels.txt gpurun_out/r2_ncu_traffic.json | tail -18
LAYERS=80 ITERS=1 UMB_PDL=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/ncu_iteration.py > gpurun_out/r2_ncu_iter.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches.csv 0 40 > gpurun_out/r2_launches_iteration.txt; head -30 gpurun_out/r2_launches_iteration.txt

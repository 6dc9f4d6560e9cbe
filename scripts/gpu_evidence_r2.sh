#!/bin/bash
# round-2 evidence (files): big artefacts stay in /tmp on the box, only summaries go to gpurun_out/ (64 MiB pull limit)
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 64 --warmup 5 --with-offload > gpurun_out/r2_bench_static70b_final.json 2> gpurun_out/r2_bench_final.err
echo "static70b rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_static70b_final.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['verify_pass'])"; tail -3 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --workload static70b_b200tree --steps 64 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2_bench_static70b_b200tree.json 2>/dev/null
echo "b200tree rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_static70b_b200tree.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['config']['mean_accept_len'], d['config']['expected_accept_len'], d['verify_pass'])"
timeout 900 python bench.py --workload stochastic32k_hbm --steps 32 --warmup 3 --no-reference-gpu > gpurun_out/r2_bench_stochastic32k_hbm.json 2>/dev/null
echo "stochastic32k_hbm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -o /tmp/r2_kernels_full -f python scripts/ncu_gemm.py > gpurun_out/r2_ncu_kernels.log 2>&1
python scripts/ncu_traffic.py /tmp/r2_kernels_full.ncu-rep gpurun_out/r2_ncu_kernels.txt gpurun_out/r2_ncu_traffic.json | grep -c gemm_kernel
ls -la /tmp/r2_kernels_full.ncu-rep
LAYERS=80 ITERS=1 UMB_PDL=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/r2_launches.csv python scripts/ncu_iteration.py > gpurun_out/r2_ncu_iter.log 2>&1
python scripts/launch_summary.py /tmp/r2_launches.csv 0 60 > gpurun_out/r2_launches_iteration.txt; head -3 gpurun_out/r2_launches_iteration.txt
du -sh gpurun_out

#!/bin/bash
bash scripts/gpu_final.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|tree_attn|tree_accept|kv_gather|rmsnorm|rope_append' -o gpurun_out/kernels_full_final -f python scripts/ncu_gemm.py > gpurun_out/ncu_kernels_final.log 2>&1; tail -1 gpurun_out/ncu_kernels_final.log
UMB_PDL=0 LAYERS=80 ITERS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python scripts/ncu_iteration.py > gpurun_out/ncu_iter_final.log 2>&1; tail -1 gpurun_out/ncu_iter_final.log

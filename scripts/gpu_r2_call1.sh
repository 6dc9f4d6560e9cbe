#!/bin/bash
# round 2, call 1: tcgen05 rate microbenchmarks, the reference-kernel probe, the whole GPU suite incl. the 70B parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt
timeout 120 scripts/ubench/tc_rates > gpurun_out/r2_tc_rates.txt 2>&1
echo "tc_rates rc=$?"; cat gpurun_out/r2_tc_rates.txt
timeout 900 python scripts/probe_ref_kernels.py > gpurun_out/r2_probe_ref.txt 2>&1
echo "probe rc=$?"; tail -20 gpurun_out/r2_probe_ref.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/r2_pytest1.txt 2>&1
echo "pytest rc=$?"; grep -E "parity70b|passed|failed|Error|error" gpurun_out/r2_pytest1.txt | tail -30

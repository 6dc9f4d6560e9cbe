#!/bin/bash
# profiling bundle: GEMM timeline, stage sweep, full ncu of GEMM + attention, launch list, sanitizer
mkdir -p gpurun_out
timeout 120 python scripts/gemm_trace.py > gpurun_out/gemm_trace.log 2>&1
: > gpurun_out/sweep_stages.log
for st in 6 8 10 12 14 16; do
  UMB_GEMM_STAGES=$st timeout 120 python scripts/bench_gemm_dbg.py >> gpurun_out/sweep_stages.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 3 -o gpurun_out/gemm_full -f python scripts/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_attn -c 3 -o gpurun_out/attn_full -f python scripts/ncu_gemm.py > gpurun_out/ncu_attn.log 2>&1
UMB_PDL=0 LAYERS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/ncu_iteration.py > gpurun_out/ncu_iter.log 2>&1
tail -3 gpurun_out/ncu_iter.log
LAYERS=1 PROMPT=160 ITERS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/ncu_iteration.py > gpurun_out/sanitizer.log 2>&1
echo sanitizer rc=$?; tail -5 gpurun_out/sanitizer.log

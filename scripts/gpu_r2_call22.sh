#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python scripts/draft_forward.py 2>&1 | tail -2
UMB_SKINNY=0 timeout 300 python scripts/draft_forward.py 2>&1 | tail -2
UMB_PDL=0 timeout 300 python scripts/draft_forward.py 2>&1 | tail -2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_draft_launches.csv python scripts/draft_forward.py 2>&1 | tail -2
python scripts/launch_summary.py gpurun_out/r2_draft_launches.csv 0 40
UMB_SKINNY=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_draft_launches_noskinny.csv python scripts/draft_forward.py 2>&1 | tail -2
python scripts/launch_summary.py gpurun_out/r2_draft_launches_noskinny.csv 0 40
for v in xprod xprod4; do
  export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/bench_kernels.py 2>&1 | grep -E "awq_gemm|4 GEMMs|---|rror" | head -12
done
} > gpurun_out/r2_call22.txt 2>&1
cat gpurun_out/r2_call22.txt

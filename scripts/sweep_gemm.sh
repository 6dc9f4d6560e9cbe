#!/bin/bash
# sweep the GEMM pipeline knobs; one process per configuration
out=gpurun_out/sweep_gemm.log; : > $out
for pf in 0 8 16 32 64; do for st in 10 14 16; do
  UMB_GEMM_PREFETCH=$pf UMB_GEMM_STAGES=$st timeout 120 python scripts/bench_gemm_dbg.py >> $out 2>&1
done; done
UMB_GEMM_PREFETCH=32 F16=1 timeout 120 python scripts/bench_gemm_dbg.py >> $out 2>&1
UMB_GEMM_PREFETCH=0 F16=1 timeout 120 python scripts/bench_gemm_dbg.py >> $out 2>&1
grep -E "config|layer sum|gemm_f16" $out

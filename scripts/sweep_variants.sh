#!/bin/bash
# one bench_gemm_dbg.py process per experiment library (umbrella_b200/build.py --variant NAME -D...)
out=gpurun_out/sweep_variants.log; : > $out
for v in "$@"; do
  echo "=== variant $v" >> $out
  UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.$v.so timeout 120 python scripts/bench_gemm_dbg.py >> $out 2>&1
done
grep -E "variant|layer sum|awq_gemm" $out

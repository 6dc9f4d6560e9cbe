#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python scripts/iter_breakdown.py > gpurun_out/iter_fused$i.log 2>&1; echo "fused $i rc=$?"; grep -E "graph=True|draft levels  |target verify forward  " gpurun_out/iter_fused$i.log | tail -3
  UMB_FUSE_ROPE=0 timeout 300 python scripts/iter_breakdown.py > gpurun_out/iter_unfused$i.log 2>&1; echo "unfused $i rc=$?"; grep -E "graph=True|draft levels  |target verify forward  " gpurun_out/iter_unfused$i.log | tail -3
done

#!/bin/bash
mkdir -p gpurun_out
for cfg in "UMB_PDL=1 GRAPH=1" "UMB_PDL=1 GRAPH=0 DECODE=0" "UMB_PDL=0 GRAPH=0 DECODE=0" "UMB_PDL=1 GRAPH=1"; do
  env $cfg REPS=40 timeout 400 python scripts/stress_prefill.py 2>&1 | grep -E "FAILED|stress done" | tail -2
done

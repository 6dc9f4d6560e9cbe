#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gemm_unit_trace.py > gpurun_out/r2_unit_trace_drain3.txt 2>&1
tail -3 gpurun_out/r2_unit_trace_drain3.txt

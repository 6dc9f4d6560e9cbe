#!/bin/bash
mkdir -p gpurun_out
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tcdbg.so
timeout 240 python scripts/one_attn.py 8 1 128 4100 2>&1 | grep -v "^$" | sort | uniq -c | sort -rn | head -30

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/tests26.log 2>&1; tail -3 gpurun_out/tests26.log
timeout 300 python scripts/iter_breakdown.py 2>&1 | grep -E "graph=|graph replay"
UMB_FUSE_ROPE=0 timeout 300 python scripts/iter_breakdown.py 2>&1 | grep -E "graph=True|graph replay"

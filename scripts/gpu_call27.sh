#!/bin/bash
python -m pytest tests/test_gpu_glue.py tests/test_gpu_attn.py tests/test_gpu_forward.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -2
for t in 32 16 8; do echo "== slab tiles $t"; UMB_ATTN_SLAB_TILES=$t timeout 300 python scripts/iter_breakdown.py 2>&1 | grep -E "graph=True|graph replay"; done

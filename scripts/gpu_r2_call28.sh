#!/bin/bash
mkdir -p gpurun_out
{
echo "--- draft forward knobs"
timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
UMB_FUSE_ROPE=1 timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
UMB_GEMM_ALIGNED_SPLIT=0 timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
UMB_GEMM_STAGES=4 timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
UMB_ATTN_SLAB_TILES=4 timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
ROWS=2 timeout 300 python scripts/draft_forward.py 2>&1 | grep "draft forward"
} > gpurun_out/r2_draft_knobs.txt 2>&1
cat gpurun_out/r2_draft_knobs.txt
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.m128.so
M=128 REPS=3000 timeout 900 cuda-gdb -batch -ex "set pagination off" -ex "set cuda break_on_launch none" -ex run -ex "info cuda kernels" -ex bt -ex "x/10i \$pc-64" -ex "info cuda warps" --args python scripts/repro_m128.py > gpurun_out/r2_m128_gdb.txt 2>&1
grep -v "^\[New Thread\|^\[Thread\|warning:" gpurun_out/r2_m128_gdb.txt | head -120

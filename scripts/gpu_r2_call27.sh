#!/bin/bash
mkdir -p gpurun_out
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.m128.so
export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_COREDUMP_FILE=/tmp/m128_core CUDA_COREDUMP_GENERATION_FLAGS=skip_global_memory,skip_local_memory,skip_constbank_memory
{
M=128 REPS=2000 timeout 900 python scripts/repro_m128.py 2>&1 | tail -4
ls -la /tmp/m128_core* 2>&1
cat > /tmp/gdbcmds <<'G'
set pagination off
target cudacore /tmp/m128_core
info cuda kernels
bt
x/12i $pc-48
info cuda warps
info cuda lanes
G
timeout 600 cuda-gdb -batch -x /tmp/gdbcmds 2>&1 | head -150
} > gpurun_out/r2_m128_core.txt 2>&1
cat gpurun_out/r2_m128_core.txt | head -120
unset UMB_LIB_PATH CUDA_ENABLE_COREDUMP_ON_EXCEPTION
{
echo "--- draft forward knobs"
timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
UMB_FUSE_ROPE=1 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
UMB_GEMM_ALIGNED_SPLIT=0 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
UMB_GEMM_STAGES=4 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
UMB_ATTN_SLAB_TILES=4 timeout 300 python scripts/draft_forward.py 2>&1 | tail -1
} > gpurun_out/r2_draft_knobs.txt 2>&1
cat gpurun_out/r2_draft_knobs.txt

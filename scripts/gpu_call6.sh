#!/bin/bash
# evidence bundle: all GPU tests, bench (both arms), full ncu of every hot kernel, whole-iteration launch list
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_tree.py tests/test_gpu_glue.py tests/test_gpu_attn.py tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_dynamic.py tests/test_gpu_offload.py > gpurun_out/tests6.log 2>&1
grep -E "^==|passed|failed" gpurun_out/tests6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 300 gpurun_out/bench6.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench6_ref.json 2> gpurun_out/bench6_ref.err; tail -c 300 gpurun_out/bench6_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|tree_attn|tree_accept|kv_gather' -o gpurun_out/kernels_full -f python scripts/ncu_gemm.py > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log
UMB_PDL=0 LAYERS=80 ITERS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv python scripts/ncu_iteration.py > gpurun_out/ncu_iter6.log 2>&1; tail -2 gpurun_out/ncu_iter6.log

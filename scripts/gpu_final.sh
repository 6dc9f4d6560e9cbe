#!/bin/bash
# round-end evidence: every GPU test, smoke, both bench arms, kernel micro-bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/tests_final.log 2>&1; tail -3 gpurun_out/tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 200 gpurun_out/bench_final_ref.json
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2600 gpurun_out/bench_final.json | head -c 1400
timeout 200 python scripts/bench_kernels.py > gpurun_out/kernels_final.log 2>&1; grep -E "layer|M = 13" -A5 gpurun_out/kernels_final.log | head -8

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "n2 rc=$?"; python -c "import json; d=json.loads([l for l in open('gpurun_out/r2_bench_n2.json') if l.startswith('{')][-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], d['replicas'])"; tail -3 gpurun_out/r2_bench_n2.err
timeout 600 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

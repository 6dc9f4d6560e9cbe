#!/bin/bash
# Run every GPU test file in its own process under a timeout (a hung kernel must not eat the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
rc=0
for f in ${@:-tests/test_gpu_tree.py tests/test_gpu_glue.py tests/test_gpu_attn.py tests/test_gpu_gemm.py tests/test_gpu_forward.py}; do
  name=$(basename $f .py)
  timeout -s KILL 600 python -m pytest $f -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f rc=$r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=1
done
exit $rc

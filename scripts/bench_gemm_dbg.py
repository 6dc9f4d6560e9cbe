"""One line per GEMM shape of the bench workload at M rows (env M, default 13); used by the shell sweeps
over UMB_GEMM_STAGES / UMB_GEMM_PREFETCH / UMB_LIB_PATH (the env is read once per process)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
from scripts.bench_kernels import bench_awq, bench_f16
M = int(os.environ.get("M", "13"))
print("config:", {k: v for k, v in os.environ.items() if k.startswith("UMB_")}, "M", M)
t = [bench_awq(M, 8192, 10240), bench_awq(M, 8192, 8192, ops.UMB_EPI_RESIDUAL),
     bench_awq(M, 8192, 57344, ops.UMB_EPI_SILU_MUL), bench_awq(M, 28672, 8192, ops.UMB_EPI_RESIDUAL)]
print(f"    layer sum: {sum(t)*1e3:.1f} us")
if os.environ.get("F16", "0") == "1":
    # Llama-3.2-1B draft shapes (fp16 weights) at 3 rows
    for (k, n) in ((2048, 3072), (2048, 2048), (2048, 16384), (8192, 2048)):
        bench_f16(3, k, n, copies=64)

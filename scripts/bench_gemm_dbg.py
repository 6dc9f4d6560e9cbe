import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
from scripts.bench_kernels import bench_awq
bench_awq(13, 8192, 57344, ops.UMB_EPI_SILU_MUL)
bench_awq(13, 28672, 8192, ops.UMB_EPI_RESIDUAL)

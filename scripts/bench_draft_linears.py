"""fp16 linears of the draft model (Llama-3.2-1B shapes) at a few token rows: CUDA events, weights rotated through
> 126 MB.  UMB_SKINNY=0 times the tensor-memory kernel on the same shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops  # noqa: E402

DEV = "cuda:0"


def time_fn(fn, iters=50, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def bench(M, K, N, mode):
    copies = int(300e6 // (N * K * 2)) + 2
    ws_ = [(0.02 * torch.randn(N, K, device=DEV)).to(torch.float16) for _ in range(copies)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    n_out = N // 2 if mode == ops.UMB_EPI_SILU_MUL else N
    out = torch.zeros(M, n_out, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    ms = time_fn(lambda i: ops.gemm_f16(x, ws_[i % copies], **kw))
    gbs = N * K * 2 / ms / 1e6
    print(f"gemm_f16 M={M} K={K:5d} N={N:6d} mode={mode}: {ms * 1e3:7.1f} us  {gbs:7.1f} GB/s")
    return ms


if __name__ == "__main__":
    print("UMB_SKINNY =", os.environ.get("UMB_SKINNY", "1"))
    for M in (2, 4, 8):
        tot = bench(M, 2048, 3072, ops.UMB_EPI_STORE) + bench(M, 2048, 2048, ops.UMB_EPI_RESIDUAL) + \
            bench(M, 2048, 16384, ops.UMB_EPI_SILU_MUL) + bench(M, 8192, 2048, ops.UMB_EPI_RESIDUAL)
        print(f"  4 linears of one 1B layer at M={M}: {tot * 1e3:.1f} us (HBM roofline 18.6 us)")

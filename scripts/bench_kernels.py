"""Micro-benchmarks of the hot kernels at Llama-3.1-70B shapes (CUDA events, L2 flushed by rotating
through > 126 MB of distinct weights).  Prints achieved GB/s against the measured HBM peak."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops  # noqa: E402

DEV = "cuda:0"
PEAK = 6542.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def time_fn(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def bench_awq(M, K, N, mode=ops.UMB_EPI_STORE, copies=8):
    nbytes = ops.awq_packed_bytes(K, N)
    copies = max(copies, int(400e6 // nbytes) + 1)
    packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(copies)]
    for p in packed:   # sane fp16 scales: overwrite scale bytes region pattern is irrelevant for timing
        pass
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    n_out = N // 2 if mode == ops.UMB_EPI_SILU_MUL else N
    out = torch.zeros(M, n_out, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    ms = time_fn(lambda i: ops.awq_gemm(x, packed[i % copies], K, N, **kw))
    gbs = nbytes / ms / 1e6
    print(f"awq_gemm M={M:4d} K={K:6d} N={N:6d} mode={mode}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s  {gbs/PEAK:5.2f} of measured peak")
    return ms


def bench_f16(M, K, N, copies=2):
    w = [(0.02 * torch.randn(N, K, device=DEV)).to(torch.float16) for _ in range(copies)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    ms = time_fn(lambda i: ops.gemm_f16(x, w[i % copies], mode=ops.UMB_EPI_LOGITS, out=out, workspace=ws))
    gbs = N * K * 2 / ms / 1e6
    print(f"gemm_f16 M={M:4d} K={K:6d} N={N:6d} logits: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s  {gbs/PEAK:5.2f} of measured peak")
    return ms


def bench_attn(qo, prefix, Hq=64, Hkv=8, D=128):
    kv = prefix + qo
    q = torch.randn(qo, Hq, D, device=DEV).to(torch.float16)
    k = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
    v = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
    bits = torch.full((qo, (qo + 31) // 32), -1, dtype=torch.int32, device=DEV)
    out = torch.empty_like(q)
    ws = ops.tree_attn_workspace(qo, Hq, D, DEV)
    m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
    ms = time_fn(lambda i: ops.tree_attn(q, k, v, kv, m, out=out, workspace=ws), iters=50)
    gbs = kv * Hkv * D * 2 * 2 / ms / 1e6
    print(f"tree_attn qo={qo:4d} kv={kv:6d}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s (KV bytes)")
    return ms


if __name__ == "__main__":
    t0 = time.time()
    total = 0.0
    for M in (13, 31, 64):
        print(f"--- M = {M}")
        qkv = bench_awq(M, 8192, 10240)
        o = bench_awq(M, 8192, 8192, ops.UMB_EPI_RESIDUAL)
        gu = bench_awq(M, 8192, 57344, ops.UMB_EPI_SILU_MUL)
        dn = bench_awq(M, 28672, 8192, ops.UMB_EPI_RESIDUAL)
        layer = qkv + o + gu + dn
        print(f"    4 GEMMs of one 70B layer: {layer*1e3:.1f} us -> x80 = {layer*80:.2f} ms (HBM roofline 5.44 ms)")
    bench_f16(13, 8192, 128256)
    for prefix in (512, 2048, 8192, 32768):
        bench_attn(13, prefix)
    print(f"done in {time.time()-t0:.1f}s")

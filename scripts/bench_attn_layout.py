"""Tree attention at long context, NHD [S, Hkv, D] against HND [Hkv, S, D] cache layout (CUDA events, mean of 50)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
DEV = "cuda:0"
def run(qo, prefix, layout, Hq=64, Hkv=8, D=128):
    kv = prefix + qo
    q = torch.randn(qo, Hq, D, device=DEV).half()
    shape = (kv, Hkv, D) if layout == "NHD" else (Hkv, kv, D)
    k, v = torch.randn(shape, device=DEV).half(), torch.randn(shape, device=DEV).half()
    bits = torch.full((qo, 1), -1, dtype=torch.int32, device=DEV)
    out = torch.empty_like(q)
    ws = ops.tree_attn_workspace(qo, Hq, D, DEV)
    m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
    f = lambda: ops.tree_attn(q, k, v, kv, m, out=out, workspace=ws, layout=layout)
    for _ in range(5): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50): f()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 50 * 1e3
    print(f"tree_attn qo={qo} kv={kv} {layout}: {us:7.1f} us  {kv * Hkv * D * 4 / us / 1e3:7.1f} GB/s", flush=True)
for layout in ("NHD", "HND"):
    for prefix in (8192, 32768):
        run(13, prefix, layout)

"""Tree attention (13 tokens x 64 heads, head_dim 128, NHD) over context lengths: CUDA events, mean of 50."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
DEV = "cuda:0"
for prefix in (512, 1024, 2048, 4096, 6144, 8192, 16384):
    qo, Hq, Hkv, D = 13, 64, 8, 128
    kv = prefix + qo
    q = torch.randn(qo, Hq, D, device=DEV).half()
    k, v = torch.randn(kv, Hkv, D, device=DEV).half(), torch.randn(kv, Hkv, D, device=DEV).half()
    bits = torch.full((qo, 1), -1, dtype=torch.int32, device=DEV)
    out = torch.empty_like(q)
    ws = ops.tree_attn_workspace(qo, Hq, D, DEV)
    m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
    f = lambda: ops.tree_attn(q, k, v, kv, m, out=out, workspace=ws)
    for _ in range(5): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50): f()
    e.record(); torch.cuda.synchronize()
    print(f"kv={kv:6d}: {s.elapsed_time(e) / 50 * 1e3:7.1f} us", flush=True)

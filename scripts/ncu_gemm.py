"""Tiny driver for ncu: a few launches of the AWQ GEMM (70B gate|up shape) and the tree attention."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
DEV = "cuda:0"
M = int(os.environ.get("M", "13"))
K, N = 8192, 57344
nbytes = ops.awq_packed_bytes(K, N)
packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(2)]
x = torch.randn(M, K, device=DEV).to(torch.float16)
out = torch.zeros(M, N // 2, dtype=torch.float16, device=DEV)
ws = ops.gemm_workspace(M, K, N, DEV)
for i in range(4):
    ops.awq_gemm(x, packed[i % 2], K, N, mode=ops.UMB_EPI_SILU_MUL, out=out, workspace=ws)
qo, prefix, Hq, Hkv, D = 13, 512, 64, 8, 128
kv = prefix + qo
q = torch.randn(qo, Hq, D, device=DEV).to(torch.float16)
k = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
v = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
bits = torch.full((qo, 1), -1, dtype=torch.int32, device=DEV)
o = torch.empty_like(q)
wsa = ops.tree_attn_workspace(qo, Hq, D, DEV)
m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
for i in range(3):
    ops.tree_attn(q, k, v, kv, m, out=o, workspace=wsa)
torch.cuda.synchronize()

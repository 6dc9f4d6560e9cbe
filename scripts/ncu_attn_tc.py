import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
DEV = "cuda:0"
qo, prefix, Hq, Hkv, D = 13, 32768, 64, 8, 128
kv = prefix + qo
q = torch.randn(qo, Hq, D, device=DEV).half()
k, v = torch.randn(kv, Hkv, D, device=DEV).half(), torch.randn(kv, Hkv, D, device=DEV).half()
bits = torch.full((qo, 1), -1, dtype=torch.int32, device=DEV)
out = torch.empty_like(q)
ws = ops.tree_attn_workspace(qo, Hq, D, DEV)
m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
for _ in range(3):
    ops.tree_attn(q, k, v, kv, m, out=out, workspace=ws)
torch.cuda.synchronize()

"""Tiny driver for ncu: the four AWQ GEMMs of one Llama-3.1-70B layer at M rows (default 13), the fp16 lm_head
GEMM, the tree attention, the accept kernel and the KV gather -- two launches each (the second is profiled warm).

    ncu --set full --clock-control none --import-source on -o gpurun_out/kernels_full -f python scripts/ncu_gemm.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
DEV = "cuda:0"
M = int(os.environ.get("M", "13"))
for (K, N, mode) in ((8192, 10240, ops.UMB_EPI_STORE), (8192, 8192, ops.UMB_EPI_RESIDUAL),
                     (8192, 57344, ops.UMB_EPI_SILU_MUL), (28672, 8192, ops.UMB_EPI_RESIDUAL)):
    nbytes = ops.awq_packed_bytes(K, N)
    packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(2)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N // 2 if mode == ops.UMB_EPI_SILU_MUL else N, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    for i in range(2):
        ops.awq_gemm(x, packed[i % 2], K, N, **kw)
    torch.cuda.synchronize()
    del packed
# lm_head (fp16 weights, fp32 logits + arg-max keys)
V, H = 128256, 8192
w = (0.02 * torch.randn(V, H, device=DEV)).to(torch.float16)
x = torch.randn(M, H, device=DEV).to(torch.float16)
keys = torch.zeros(M, dtype=torch.int64, device=DEV)
lg = torch.empty(M, V, dtype=torch.float32, device=DEV)
ws = ops.gemm_workspace(M, H, V, DEV)
for i in range(2):
    ops.gemm_f16(x, w, mode=ops.UMB_EPI_LOGITS, argmax=keys, out=lg, workspace=ws)
del w
# tree attention, 70B head geometry, prefix 512
qo, prefix, Hq, Hkv, D = M, 512, 64, 8, 128
kv = prefix + qo
q = torch.randn(qo, Hq, D, device=DEV).to(torch.float16)
k = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
v = torch.randn(kv, Hkv, D, device=DEV).to(torch.float16)
bits = torch.full((qo, 1), -1, dtype=torch.int32, device=DEV)
o = torch.empty_like(q)
wsa = ops.tree_attn_workspace(qo, Hq, D, DEV)
m = ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits)
for i in range(2):
    ops.tree_attn(q, k, v, kv, m, out=o, workspace=wsa)
# accept + gather (3x4 Sequoia tree, 80-layer 70B cache of 2048 rows)
from umbrella_b200.sequoia_utils import load_growmap
gm = load_growmap("sequoia_tree-3x4.json")
N = gm["size"]
parents = torch.zeros(N, dtype=torch.int32)
for vtx, succ in enumerate(gm["Successors"]):
    for c in succ:
        parents[c] = vtx
parents = parents.to(DEV)
depth = torch.tensor(gm["depth"], dtype=torch.int32, device=DEV)
tokens = torch.randint(3, 30000, (2048,), device=DEV)
pos = torch.arange(2048, device=DEV)
sampled = tokens[prefix + 1:prefix + 1 + N].clone()          # accept a chain
result = torch.zeros(4 + N, dtype=torch.int32, device=DEV)
kc = torch.zeros(80, 2048, 8, 128, dtype=torch.float16, device=DEV)
vc = torch.zeros_like(kc)
for i in range(2):
    ops.tree_accept(sampled, tokens, pos, parents, depth, prefix, None, result)
    ops.kv_gather(kc, vc, result, prefix)
torch.cuda.synchronize()
print("accept_len", int(result[0]))

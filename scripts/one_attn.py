"""One long-context tree attention call against the float64 oracle (debug helper for the tcgen05 kernel)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import attention as oattn
from oracle import tree as otree
from umbrella_b200 import ops
Hq, Hkv, D, prefix = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 1, 128, 4100)))
trees = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trees.json")))["shipped"]
tm = np.asarray(otree.growmap_from_successors(trees["sequoia_tree-3x4.json"]["Successors"])["mask"], dtype=bool)
N = tm.shape[0]
rng = np.random.default_rng(11)
kv = prefix + N
q = rng.standard_normal((N, Hq, D)).astype(np.float16)
k = rng.standard_normal((kv + 7, Hkv, D)).astype(np.float16)
v = rng.standard_normal((kv + 7, Hkv, D)).astype(np.float16)
ref = oattn.tree_attention(q, k[:kv], v[:kv], prefix, tm)
qd, kd, vd = (torch.from_numpy(a).cuda() for a in (q, k, v))
bits = torch.from_numpy(otree.tree_bits_from_mask(tm).view(np.int32)).cuda()
out = ops.tree_attn(qd, kd, vd, kv, ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits))
torch.cuda.synchronize()
err = np.abs(out.cpu().numpy().astype(np.float64) - ref).max()
print(f"Hq={Hq} Hkv={Hkv} prefix={prefix}: max abs err {err:.3e} (ref max {np.abs(ref).max():.2f})")

"""Launch the (unfinished) tcgen05 attention once WITHOUT synchronising and read its progress marks from pinned memory."""
import ctypes as C, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tree as otree
from umbrella_b200 import ops, _lib
Hq, Hkv, D, prefix = 64, 8, 128, 32768
trees = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trees.json")))["shipped"]
tm = np.asarray(otree.growmap_from_successors(trees["sequoia_tree-3x4.json"]["Successors"])["mask"], dtype=bool)
N = tm.shape[0]
kv = prefix + N
qd = torch.randn(N, Hq, D, device="cuda").half()
kd = torch.randn(kv + 7, Hkv, D, device="cuda").half()
vd = torch.randn(kv + 7, Hkv, D, device="cuda").half()
bits = torch.from_numpy(otree.tree_bits_from_mask(tm).view(np.int32)).cuda()
dbg = torch.zeros(256, dtype=torch.int64).pin_memory()
lib = _lib.load()
lib.umb_debug_attn_tc_ptr.argtypes = [C.c_void_p]
print("set ptr rc", lib.umb_debug_attn_tc_ptr(C.c_void_p(dbg.data_ptr())), flush=True)
torch.cuda.synchronize()
out = ops.tree_attn(qd, kd, vd, kv, ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits))
print("launched", flush=True)
out2 = ops.tree_attn(qd, kd, vd, kv, ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits))
torch.cuda.synchronize()
t0 = int(dbg[64])
print("tile: softmax[wait_s, s_ready, p_stored, o_ready, p_signalled]  mma[loop, p_full seen, pv issued]  (ns from tile 0)")
for t in range(12):
    print(f"{t:2d}: " + " ".join(f"{(int(dbg[64 + t * 8 + k]) - t0) if int(dbg[64 + t * 8 + k]) else -1:7d}" for k in range(8)), flush=True)
os._exit(0)

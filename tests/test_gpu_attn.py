"""GPU parity: tree-masked GQA attention vs the float64 oracle (oracle/attention.py, itself pinned
against the reference's StaticKV_Cache.compute_attention in tests/test_oracle_numeric.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import tree as otree

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
TOL = 2e-3      # fp16 output (ulp 4.9e-4 relative) + fp16 P in the PV product; stated vs max |out|


def _tree(name):
    trees = json.load(open(os.path.join(GOLDEN, "trees.json")))["shipped"]
    gm = otree.growmap_from_successors(trees[name]["Successors"])
    return np.asarray(gm["mask"], dtype=bool)


def _check(out, ref):
    err = np.abs(out.astype(np.float64) - ref).max()
    assert err <= TOL * max(1.0, np.abs(ref).max()), err


@pytest.mark.parametrize("tree_name,Hq,Hkv,D,prefix", [
    ("sequoia_tree-3x4.json", 64, 8, 128, 512),
    ("sequoia_tree-3x4.json", 64, 8, 128, 5),
    ("8b_sequoia_tree-5x6.json", 32, 8, 128, 300),
    ("8b_sequoia_tree-6x7.json", 32, 8, 64, 1000),
    ("sequoia_tree-3x4.json", 8, 1, 128, 4100),       # long context -> split-KV + merge (tcgen05 kernel from 4096 rows)
    ("sequoia_tree-3x4.json", 64, 8, 128, 9000),      # tcgen05 kernel, 8 kv heads x 18 splits, last tile ragged
    ("sequoia_tree-3x4.json", 32, 8, 128, 6011),      # G = 4: 52 of the 128 tile rows in use
    ("sequoia_tree-3x4.json", 64, 8, 128, 4083),      # tree rows straddle a 128-row tile boundary
])
def test_tree_mask_modes(tree_name, Hq, Hkv, D, prefix):
    from umbrella_b200 import ops
    tm = _tree(tree_name)
    N = tm.shape[0]
    rng = np.random.default_rng(11)
    kv = prefix + N
    S = kv + 7
    q = rng.standard_normal((N, Hq, D)).astype(np.float16)
    k = rng.standard_normal((S, Hkv, D)).astype(np.float16)
    v = rng.standard_normal((S, Hkv, D)).astype(np.float16)
    ref = oattn.tree_attention(q, k[:kv], v[:kv], prefix, tm)
    qd, kd, vd = (torch.from_numpy(a).to(DEV) for a in (q, k, v))
    bits = torch.from_numpy(otree.tree_bits_from_mask(tm).view(np.int32)).to(DEV)
    out = ops.tree_attn(qd, kd, vd, kv, ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits))
    _check(out.cpu().numpy(), ref)
    # same rows through the dense-mask form (what KV_Cache.compute_attention receives, cache.py:82)
    dense = torch.from_numpy(otree.dense_rows_from_compact(prefix, tm, S)).to(DEV)
    out2 = ops.tree_attn(qd, kd, vd, kv, ops.make_mask(ops.UMB_MASK_DENSE, dense_mask=dense))
    _check(out2.cpu().numpy(), ref)
    # HND cache layout (StaticKV_Cache, cache.py:98-134)
    kh, vh = kd.transpose(0, 1).contiguous(), vd.transpose(0, 1).contiguous()
    out3 = ops.tree_attn(qd, kh, vh, kv, ops.make_mask(ops.UMB_MASK_TREE, prefix, tree_bits=bits), layout="HND")
    _check(out3.cpu().numpy(), ref)


@pytest.mark.parametrize("qo,prefix,Hq,Hkv,D", [(128, 0, 8, 2, 128), (77, 130, 32, 8, 64), (1, 600, 64, 8, 128),
                                                (16, 5000, 64, 8, 128), (3, 4500, 8, 1, 128)])     # tcgen05 kernel
def test_causal(qo, prefix, Hq, Hkv, D):
    from umbrella_b200 import ops
    rng = np.random.default_rng(5)
    kv = prefix + qo
    q = rng.standard_normal((qo, Hq, D)).astype(np.float16)
    k = rng.standard_normal((kv, Hkv, D)).astype(np.float16)
    v = rng.standard_normal((kv, Hkv, D)).astype(np.float16)
    mask = np.arange(kv)[None, :] <= (prefix + np.arange(qo))[:, None]
    ref = oattn.masked_gqa_attention(q, k, v, mask)
    out = ops.tree_attn(*(torch.from_numpy(a).to(DEV) for a in (q, k, v)), kv, ops.make_mask(ops.UMB_MASK_CAUSAL, prefix))
    _check(out.cpu().numpy(), ref)

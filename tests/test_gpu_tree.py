"""GPU parity: tree accept + KV gather (integer, bit-exact) vs the oracle and the reference-generated
golden fixtures (tests/golden/accept_*.json|npz, produced by oracle/make_golden.py from the
reference's own verify())."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import tree as otree

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _dev():
    return torch.device("cuda:0")


def _run_accept(ops, sampled, tokens, position_ids, parents, depth, n, eos):
    dev = _dev()
    N = len(parents)
    t = torch.from_numpy(tokens.copy()).to(dev)
    p = torch.from_numpy(position_ids.copy()).to(dev)
    res = torch.zeros(4 + N, dtype=torch.int32, device=dev)
    e = torch.tensor(eos, dtype=torch.int64, device=dev) if len(eos) else None
    ops.tree_accept(torch.from_numpy(sampled).to(dev), t, p, torch.from_numpy(parents.astype(np.int32)).to(dev),
                    torch.from_numpy(depth.astype(np.int32)).to(dev), n, e, res)
    torch.cuda.synchronize()
    return t.cpu().numpy(), p.cpu().numpy(), res.cpu().numpy()


def _oracle_step(tokens, position_ids, n, sampled, parents, mask, depth, eos):
    st = otree.VerifyState(len(tokens), depth)
    st.tokens = tokens.copy()
    st.position_ids = position_ids.copy()
    st.num_nodes = n
    cont, path, alen = otree.verify_step(st, sampled, parents, mask, mask.sum(-1).astype(np.int32), eos)
    return st, cont, path, alen


@pytest.mark.parametrize("tree_name", ["sequoia_tree-3x4.json", "8b_sequoia_tree-5x6.json", "8b_sequoia_tree-6x7.json"])
def test_accept_random_vs_oracle(tree_name):
    from umbrella_b200 import ops
    trees = json.load(open(os.path.join(GOLDEN, "trees.json")))["shipped"]
    gm = otree.growmap_from_successors(trees[tree_name]["Successors"] if isinstance(trees[tree_name], dict) else trees[tree_name])
    succ = gm["Successors"]
    parents = otree.successors_to_parents(succ)
    depth = np.asarray(gm["depth"], dtype=np.int64)
    mask = np.asarray(gm["mask"], dtype=bool)
    N = gm["size"]
    rng = np.random.default_rng(7)
    S = 512
    for it in range(40):
        n = int(rng.integers(1, 300))
        tokens = rng.integers(3, 1000, size=S).astype(np.int64)
        pos = np.zeros(S, dtype=np.int64)
        pos[:n] = np.arange(n)
        pos[n:n + N] = n + depth
        sampled = otree.scripted_samples(succ, tokens[n:n + N], rng, vocab_hi=1000)
        eos = []
        if it % 3 == 1:
            eos = [int(sampled[rng.integers(0, N)]), 2]
        if it % 7 == 3:
            eos = [int(tokens[n])]           # EOS at the root -> accept_length 0
        st, cont, path, alen = _oracle_step(tokens, pos, n, sampled, parents, mask, depth, eos)
        t, p, res = _run_accept(ops, sampled, tokens, pos, parents, depth, n, eos)
        assert res[0] == alen and res[1] == int(cont) and res[2] == st.num_nodes
        assert np.array_equal(res[4:4 + alen], path)
        assert np.array_equal(t, st.tokens)
        assert np.array_equal(p, st.position_ids)


def test_accept_golden_static():
    """Replay the reference-generated static-engine traces (the reference's own verify() on CPU,
    oracle/make_golden.py): tokens / position_ids / num_nodes / KV digests must match bit for bit,
    with the state carried on the GPU from iteration to iteration."""
    from umbrella_b200 import ops
    dev = _dev()
    cases = json.load(open(os.path.join(GOLDEN, "accept_static.json")))
    n_iter = 0
    for c in cases:
        parents = torch.tensor(c["parents"], dtype=torch.int32, device=dev)
        depth_np = np.asarray(c["depth"], dtype=np.int64)
        depth = torch.tensor(c["depth"], dtype=torch.int32, device=dev)
        N, S, P = len(c["parents"]), c["max_length"], len(c["prompt"])
        tokens = torch.zeros(S, dtype=torch.int64, device=dev)
        pos = torch.zeros(S, dtype=torch.int64, device=dev)
        tokens[:P] = torch.tensor(c["prompt"], device=dev)
        tokens[P] = c["first_token"]
        pos[:P] = torch.arange(P, device=dev)
        pos[P:P + N] = P + torch.from_numpy(depth_np).to(dev)
        eos = torch.tensor(c["eos"], dtype=torch.int64, device=dev) if c["eos"] else None
        # one-plane KV stand-in: row value = token * 1000 + position (as oracle/make_golden.py digests it)
        k = torch.zeros(2, S, 1, 8, dtype=torch.float16, device=dev)
        v = torch.zeros_like(k)
        res = torch.zeros(4 + N, dtype=torch.int32, device=dev)
        n = P
        for it in c["iters"]:
            assert n == it["num_nodes_before"]
            tokens[n:n + N] = torch.tensor(it["speculated"], device=dev)
            assert pos[:n + N].tolist() == it["position_ids_before"]
            k[:, n:n + N, 0, 0] = torch.arange(n, n + N, device=dev).to(torch.float16)   # tag rows by cache slot
            ops.tree_accept(torch.tensor(it["sampled"], dtype=torch.int64, device=dev), tokens, pos, parents, depth, n,
                            eos, res)
            ops.kv_gather(k, v, res, n, "LSHD", zero_upto=S)
            r = res.cpu().numpy()
            assert bool(r[1]) == it["continue"]
            assert r[2] == it["num_nodes_after"]
            n1 = int(r[2])
            assert tokens[:n1 + 1].tolist() == it["tokens_after"]
            assert pos[:n1 + N].tolist() == it["position_ids_after"]
            # gathered rows: slot n + j now holds the row that sat at path[j]
            assert k[0, n:n1, 0, 0].tolist() == [float(x) for x in r[4:4 + r[0]]]
            assert float(k[:, n1:].abs().max()) == 0.0
            n = n1
            n_iter += 1
    assert n_iter == 90


@pytest.mark.parametrize("layout", ["LSHD", "LHSD"])
def test_kv_gather(layout):
    from umbrella_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(3)
    L, S, Hkv, D = 5, 96, 4, 64
    shape = (L, S, Hkv, D) if layout == "LSHD" else (L, Hkv, S, D)
    seq_axis = 1 if layout == "LSHD" else 2
    for trial in range(12):
        k = rng.standard_normal(shape).astype(np.float16)
        v = rng.standard_normal(shape).astype(np.float16)
        offset = int(rng.integers(1, 40))
        m = int(rng.integers(0, 20))
        idx = np.sort(rng.choice(np.arange(offset, offset + 40), size=m, replace=False)).astype(np.int64)
        # includes the hazard of SURVEY 9-5 when idx[k] == offset + j for k < j
        zero_tail = trial % 2 == 0
        k_ref, v_ref = k.copy(), v.copy()
        otree.kv_gather(k_ref, idx, offset, seq_axis, zero_tail)
        otree.kv_gather(v_ref, idx, offset, seq_axis, zero_tail)
        res = torch.zeros(4 + 64, dtype=torch.int32, device=dev)
        res[0] = m
        res[4:4 + m] = torch.from_numpy(idx.astype(np.int32)).to(dev)
        kd, vd = torch.from_numpy(k).to(dev), torch.from_numpy(v).to(dev)
        ops.kv_gather(kd, vd, res, offset, layout, zero_upto=S if zero_tail else 0)
        torch.cuda.synchronize()
        assert np.array_equal(kd.cpu().numpy().view(np.uint16), k_ref.view(np.uint16))
        assert np.array_equal(vd.cpu().numpy().view(np.uint16), v_ref.view(np.uint16))

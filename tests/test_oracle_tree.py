"""Oracle (oracle/tree.py) pinned against fixtures produced by the reference's own Python
(oracle/make_golden.py): growmaps, accept, state advance, KV gather, mask windows."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import tree as T


def canon_hash(obj) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


@pytest.fixture(scope="module")
def trees(golden_dir):
    return json.load(open(os.path.join(golden_dir, "trees.json")))


def test_growmap_rebuild_matches_shipped_json(trees):
    assert len(trees["shipped"]) == 6
    for name, ref in trees["shipped"].items():
        g = T.growmap_from_successors(ref["Successors"])
        assert g["size"] == ref["size"]
        assert g["depth"] == ref["depth"]
        assert g["roots"] == ref["roots"]
        assert g["branches"] == ref["branches"]
        assert canon_hash(g["mask"]) == ref["mask_sha256"], name
        assert canon_hash(g) == ref["all_sha256"], name


def test_generate_sequoia_tree_matches_reference(trees):
    for key, ref in trees["generated"].items():
        if key == "5x5_acc":
            g = T.generate_sequoia_tree(5, 5, acc=ref["acc"])
        else:
            w, d = map(int, key.split("x"))
            g = T.generate_sequoia_tree(w, d)
        assert g["Successors"] == ref["Successors"], key
        assert canon_hash(g) == ref["all_sha256"], key
    # the shipped 3x4 growmap is exactly generate_sequoia_tree(3, 4)
    assert canon_hash(T.generate_sequoia_tree(3, 4)) == trees["shipped"]["sequoia_tree-3x4.json"]["all_sha256"]


def test_tree_invariants(trees):
    for name, ref in trees["shipped"].items():
        succ = ref["Successors"]
        parents = T.successors_to_parents(succ)
        mask = np.array(T.successors_to_mask(succ))
        depth = np.array(ref["depth"])
        assert (mask.sum(-1) == depth + 1).all()
        assert all(parents[c] < c for c in range(1, len(succ)))
        bits = T.tree_bits_from_mask(mask)
        for i in range(len(succ)):
            for j in range(len(succ)):
                assert bool((bits[i, j // 32] >> np.uint32(j % 32)) & 1) == bool(mask[i, j])


def _sim_cache_rows(tokens, positions, layers=2):
    base = np.asarray(tokens, dtype=np.float64) * 1000 + np.asarray(positions, dtype=np.float64)
    return np.stack([base + l * 0.5 for l in range(layers)])[:, :, None, None]        # [L, q, 1, 1]


def test_static_verify_replay(golden_dir, trees):
    cases = json.load(open(os.path.join(golden_dir, "accept_static.json")))
    assert len(cases) == 12
    n_iter = 0
    for c in cases:
        succ = trees["shipped"][c["tree"]]["Successors"]
        parents = T.successors_to_parents(succ)
        assert parents.tolist() == c["parents"]
        mask = np.array(T.successors_to_mask(succ)).astype(bool)
        nip = mask.sum(-1)
        assert nip.tolist() == c["node_in_path"]
        st = T.VerifyState(c["max_length"], np.array(c["depth"]))
        st.prefill(np.array(c["prompt"]), c["first_token"])
        S = c["max_length"]
        cache = np.zeros((2, S, 1, 1))
        p = len(c["prompt"])
        cache[:, :p] = _sim_cache_rows(c["prompt"], np.arange(p))
        kv_offset = p
        for it in c["iters"]:
            n0 = it["num_nodes_before"]
            assert st.num_nodes == n0
            N = st.tree_size
            st.tokens[n0:n0 + N] = it["speculated"]          # what build_tree wrote
            assert st.position_ids[:n0 + N].tolist() == it["position_ids_before"]
            # target forward appends the N tree rows at kv_offset (cache.py:60-64)
            cache[:, kv_offset:kv_offset + N] = _sim_cache_rows(it["speculated"], st.position_ids[n0:n0 + N])
            cont, path_abs, alen = T.verify_step(st, np.array(it["sampled"]), parents, mask, nip, c["eos"])
            kv_offset = T.kv_gather(cache, path_abs, n0, seq_axis=1)
            assert cont == it["continue"]
            assert st.num_nodes == it["num_nodes_after"]
            n1 = st.num_nodes
            assert st.tokens[:n1 + 1].tolist() == it["tokens_after"]
            assert st.position_ids[:n1 + N].tolist() == it["position_ids_after"]
            assert kv_offset == it["target_kv_offset"] == it["draft_kv_offset"]
            assert cache[:, :n1, 0, 0].tolist() == it["target_k_digest"]
            assert cache[:, :n1, 0, 0].tolist() == it["draft_k_digest"]
            assert it["target_tail_zero"] and (cache[:, n1:] == 0).all()
            n_iter += 1
    assert n_iter == 90


def test_scripted_accept_cases(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "accept_scripted.json")))
    succ = g["Successors"]
    parents = T.successors_to_parents(succ)
    mask = np.array(T.successors_to_mask(succ)).astype(bool)
    nip = mask.sum(-1)
    depth = nip - 1
    seen = set()
    for c in g["cases"]:
        n0, N = c["num_nodes_before"], len(succ)
        st = T.VerifyState(128, depth)
        st.num_nodes = n0
        st.tokens[:n0] = np.arange(1, n0 + 1)
        st.position_ids[:n0] = np.arange(n0)
        st.position_ids[n0:n0 + N] = n0 + depth
        st.tokens[n0:n0 + N] = c["speculated"]
        cache = np.zeros((2, 128, 1, 1))
        pos = st.position_ids[:n0 + N].astype(np.float64)
        rows = np.concatenate([pos[:n0] * 100 + np.arange(n0), pos[n0:] * 100 + np.arange(N)])
        cache[0, :n0 + N, 0, 0] = rows
        cache[1, :n0 + N, 0, 0] = rows + 1
        cont, path_abs, alen = T.verify_step(st, np.array(c["sampled"]), parents, mask, nip, c["eos"])
        off = T.kv_gather(cache, path_abs, n0)
        n1 = st.num_nodes
        assert cont == c["continue"], c["name"]
        assert n1 == c["num_nodes_after"], c["name"]
        assert st.tokens[:n1 + 1].tolist() == c["tokens_after"], c["name"]
        assert st.position_ids[:n1 + N].tolist() == c["position_ids_after"], c["name"]
        assert cache[:, :n1, 0, 0].tolist() == c["target_k_digest"], c["name"]
        assert off == c["target_kv_offset"]
        seen.add((c["name"], n1 - n0))
    assert ("eos_at_root", 0) in seen and ("root_only", 1) in seen and ("full_depth", 5) in seen


def test_dynamic_verify_replay(golden_dir):
    z = np.load(os.path.join(golden_dir, "accept_dynamic.npz"))
    for ci in range(3):
        pre = "c%d_" % ci
        meta = z[pre + "meta"]
        width, beams, dpt, vocab, S, n_it = map(int, meta[:6])
        eos = [int(e) for e in meta[6:]]
        depth = z[pre + "depth"]
        N = width * dpt + 1
        st = T.VerifyState(S, depth)
        prompt = z[pre + "prompt"]
        st.prefill(prompt, int(z[pre + "spec"][0][0]))
        cache = np.zeros((1, S, 1, 1), dtype=np.float32)
        p = len(prompt)
        cache[0, :p, 0, 0] = prompt.astype(np.float32) * 1000 + np.arange(p, dtype=np.float32)
        for it in range(n_it):
            n0 = int(z[pre + "n0"][it])
            assert st.num_nodes == n0
            spec = z[pre + "spec"][it]
            st.tokens[n0:n0 + N] = spec
            cache[0, n0:n0 + N, 0, 0] = spec.astype(np.float32) * 1000 + st.position_ids[n0:n0 + N].astype(np.float32)
            tmask = z[pre + "tree_mask"][it].astype(bool)
            parents = z[pre + "parents"][it]
            # dynamic engine: mask rows are parent row | own bit (dynamic_speculation_engine.py:247-248)
            for i in range(1, N):
                exp = tmask[parents[i]].copy(); exp[i] = True
                assert (tmask[i] == exp).all()
            cont, path_abs, alen = T.verify_step(st, z[pre + "sampled"][it], parents, tmask, depth + 1, eos)
            T.kv_gather(cache, path_abs, n0)
            n1 = st.num_nodes
            assert cont == bool(z[pre + "cont"][it]) and n1 == int(z[pre + "n1"][it])
            assert (st.tokens[:n1 + 1] == z[pre + "tokens_after"][it][:n1 + 1]).all()
            assert (st.position_ids[:n1 + N] == z[pre + "pos_after"][it][:n1 + N]).all()
            assert (cache[0, :n1, 0, 0] == z[pre + "k_digest"][it][:n1]).all()


def test_mask_windows_equal_compact_form(trees):
    """The reference's sliding window over its O(S^2) master mask equals (prefix_len, tree_mask)."""
    succ = trees["shipped"]["sequoia_tree-3x4.json"]["Successors"]
    tmask = np.array(T.successors_to_mask(succ)).astype(bool)
    N, S = len(succ), 96
    master = T.static_master_mask(S, tmask)
    for n in (1, 7, 40, S - N - 1):
        win = T.static_mask_window(master, S, n + N)
        rows = win[n:n + N]
        for kv in (n + 1, n + N):
            assert (rows[:, :kv] == T.dense_rows_from_compact(n, tmask, kv)).all()
        assert (win[:n, :n] == T.make_causal_mask(n)).all()
    dm = T.dynamic_master_mask(S, N)
    win = T.dynamic_mask_window(dm, S, 20 + N)
    assert (win[:20, :20] == T.make_causal_mask(20)).all()
    assert win[20, :21].all() and not win[21:, 21:].any()


def test_gather_overlap_hazard():
    """path [0,2,6] -> row n+2 is both a source (for slot 1) and a destination."""
    cache = np.arange(2 * 32, dtype=np.float64).reshape(2, 32, 1, 1)
    ref = cache.copy()
    n = 10
    off = T.kv_gather(cache, np.array([n, n + 2, n + 6]), n)
    assert off == n + 3
    assert (cache[:, n:n + 3] == ref[:, [n, n + 2, n + 6]]).all()
    assert (cache[:, n + 3:] == 0).all()
    c2 = ref.copy().transpose(0, 2, 1, 3).copy()           # [L, H, S, D] layout (StaticKV_Cache)
    T.kv_gather(c2, np.array([n, n + 2, n + 6]), n, seq_axis=2)
    assert (c2.transpose(0, 2, 1, 3) == cache).all()


def test_dynamic_build_tree_matches_reference(golden_dir):
    """The oracle's restatement of DynamicSpeculationEngine.build_tree reproduces the trees the reference
    built on CPU (tokens, parents, mask) from the same bigram draft table, iteration by iteration."""
    z = np.load(os.path.join(golden_dir, "accept_dynamic.npz"))
    for ci in range(3):
        pre = "c%d_" % ci
        width, beams, dpt, vocab, S, n_it = map(int, z[pre + "meta"][:6])
        N = width * dpt + 1
        table = z[pre + "draft_table"]
        for it in range(n_it):
            n0 = int(z[pre + "n0"][it])
            tokens = np.zeros(S, dtype=np.int64)
            tokens[n0] = z[pre + "spec"][it][0]                     # the root (last accepted / bonus token)
            parents, mask, _ = T.dynamic_build_tree(lambda ids: table[ids], tokens, n0, width, beams, dpt)
            assert (tokens[n0:n0 + N] == z[pre + "spec"][it]).all(), (ci, it)
            assert (parents == z[pre + "parents"][it]).all(), (ci, it)
            assert (mask == z[pre + "tree_mask"][it].astype(bool)).all(), (ci, it)

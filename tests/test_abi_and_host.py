"""CPU-side checks (no GPU, no compute calls): the C-ABI library builds / loads and exports exactly the
symbols include/umbrella_b200.h declares; the ctypes table matches the header; the host-side mirrors keep
the reference's error behaviour; the product never imports the oracle."""
import ctypes
import numpy as np
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "umbrella_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(umb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from umbrella_b200 import build, _lib
    build.build()                                       # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/umbrella_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    lib.umb_version.restype = ctypes.c_char_p
    assert lib.umb_version().startswith(b"umbrella_b200")
    lib.umb_awq_packed_bytes.restype = ctypes.c_size_t
    assert lib.umb_awq_packed_bytes(8192, 8192, 128) == 64 * 64 * 8512          # 0.5195 B / parameter
    assert lib.umb_awq_packed_bytes(8192, 8192, 64) == 0                        # unsupported group size


def test_sass_has_tcgen05_and_tma():
    """The GEMM really is a tensor-memory / TMA kernel (B200_PROFILING.md mnemonics)."""
    from umbrella_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    for mnemonic in ("UTCHMMA", "UBLKCP", "UTMALDG", "STTM", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_errors_mirror_the_reference():
    from umbrella_b200.models.auto_model import AutoModelLM
    from umbrella_b200.speculation.auto_engine import AutoEngine
    from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    with pytest.raises(ValueError):                      # auto_engine.py:14-16
        AutoEngine.from_config("cuda:0", engine="nope", model="a", draft_model="b")
    with pytest.raises(AssertionError):                  # auto_engine.py:20-21
        AutoEngine.from_config("cuda:0", engine="static", model="a")
    with pytest.raises(AssertionError):                  # static_speculation_engine.py:37
        AutoEngine.from_config("cuda:0", engine="static", model="a", draft_model="b")
    with pytest.raises(ValueError):                      # auto_model.py:166-181
        AutoModelLM.from_pretrained("unknown/model")
    e = AutoEngine.from_config("cuda:0", engine="dynamic", model="a", draft_model="b", width=4, depth=3, num_beams=6,
                               generation_length=64, template="meta-llama3")
    assert isinstance(e, DynamicSpeculationEngine) and (e.tree_width, e.tree_depth, e.num_beams) == (4, 3, 6)
    s = AutoEngine.from_config("cuda:0", engine="static", model="a", draft_model="b", growmap_path="sequoia_tree-3x4.json")
    assert isinstance(s, StaticSpeculationEngine)


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    from umbrella_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.argmax(torch.zeros(2, 8))
    for dirpath, _, files in os.walk(os.path.join(ROOT, "umbrella_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_reference_import_paths():
    """`umbrella` resolves to this package with the reference's module layout."""
    import umbrella
    from umbrella.models.auto_model import AutoModelLM                                # noqa: F401
    from umbrella.sequoia_utils import generate_sequoia_tree                            # noqa: F401
    from umbrella.speculation.auto_engine import AutoEngine                             # noqa: F401
    from umbrella.speculation.dynamic_speculation_engine import DynamicSpeculationEngine  # noqa: F401
    from umbrella.speculation.speculation_utils import (find_first_element_position, is_sentence_complete_regex,  # noqa: F401
                                                        make_causal_mask)
    from umbrella.speculation.static_speculation_engine import StaticSpeculationEngine  # noqa: F401
    import umbrella_b200
    assert umbrella.speculation.auto_engine.AutoEngine is umbrella_b200.speculation.auto_engine.AutoEngine


def test_scripted_acceptance_reproduces_expected_accept_length():
    """The benchmark's acceptance script (DEFAULT_ACC per child rank) fed through the oracle's accept block gives
    the analytic expectation of the tree (3.364 tokens per verify for sequoia_tree-3x4) -- the number bench.py
    reports as config.expected_accept_len."""
    import numpy as np
    import torch
    from oracle import tree as otree
    from umbrella_b200.sequoia_utils import expected_accept_length, load_growmap
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    gm = load_growmap("sequoia_tree-3x4.json")
    succ, N = gm["Successors"], gm["size"]
    parents = otree.successors_to_parents(succ)
    mask = np.asarray(gm["mask"], dtype=bool)
    iters = 4000
    script = ScriptedAcceptance(succ, seed=11, max_iters=iters, device="cpu")
    spec = np.arange(100, 100 + N, dtype=np.int64)              # distinct draft tokens per node
    total = 0
    for it in range(iters):
        row = script.choice[it].numpy()
        sampled = np.where(row >= 0, spec[np.clip(row, 0, None)], -1)   # -1 matches no child
        path, _ = otree.tree_accept(sampled, spec.copy(), parents, mask, mask.sum(-1))
        total += len(path)
    want = expected_accept_length(gm)
    assert abs(want - 3.3645) < 1e-3
    assert abs(total / iters - want) < 0.06


def test_product_sequoia_utils_match_reference_fixtures():
    """The PRODUCT's growmap builders (umbrella_b200/sequoia_utils.py -- what the engines and
    examples/construct_sequoia-style front-ends call) against the fixtures produced by the reference's own
    ``generate_sequoia_tree`` / shipped tree files (tests/golden/trees.json, oracle/make_golden.py)."""
    import hashlib
    import json
    from umbrella_b200.sequoia_utils import generate_sequoia_tree, growmap_from_successors, load_growmap

    def canon_hash(obj):
        return hashlib.sha256(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()

    trees = json.load(open(os.path.join(ROOT, "tests", "golden", "trees.json")))
    for key, ref in trees["generated"].items():
        if key == "5x5_acc":
            g = generate_sequoia_tree(5, 5, acc=ref["acc"])
        else:
            w, d = map(int, key.split("x"))
            g = generate_sequoia_tree(w, d)
        assert g["Successors"] == ref["Successors"], key
        assert canon_hash({k: g[k] for k in ("roots", "branches", "Successors", "mask", "depth", "size")}) == ref["all_sha256"], key
    for name, ref in trees["shipped"].items():
        gm = load_growmap(name)                                   # packaged compact form -> full schema
        assert gm["Successors"] == ref["Successors"] and gm["size"] == ref["size"], name
        assert gm["depth"] == ref["depth"] and gm["roots"] == ref["roots"] and gm["branches"] == ref["branches"], name
        assert canon_hash(gm["mask"]) == ref["mask_sha256"], name
        assert growmap_from_successors(ref["Successors"]) == gm
        assert load_growmap("../umbrella/trees/" + name) == gm    # the path form of the reference's configs
    with pytest.raises(FileNotFoundError):
        load_growmap("no_such_tree.json")


def test_engines_refuse_states_without_room_for_the_tree():
    """the kernels write tree_size rows past num_nodes (505 for the shipped dynamic config), far more
    than safe_buffer; validate_status / _prefill / _append must refuse such states on the host."""
    from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    dyn = object.__new__(DynamicSpeculationEngine)
    dyn.max_length, dyn.safe_buffer, dyn.tree_size, dyn.tree_depth = 8192, 64, 505, 24
    dyn.num_nodes = 8192 - 64                      # the reference's guard alone would let this through
    assert not dyn.validate_status()
    dyn.num_nodes = 8192 - 505 - 25
    assert dyn.validate_status()
    dyn.num_nodes += 1
    assert not dyn.validate_status()
    st = object.__new__(StaticSpeculationEngine)
    st.max_length, st.safe_buffer, st.tree_size, st.tree_depth = 2048, 16, 43, 6      # safe_buffer < tree size
    st.num_nodes = 2048 - 16
    assert not st.validate_status()
    st.num_nodes = 2048 - 43 - 7
    assert st.validate_status()
    import torch
    assert st._prefill(torch.zeros(1, 2048 - 40, dtype=torch.int64)) is False       # refused before any tensor work
    st.num_nodes = 1990
    assert st._append(torch.zeros(1, 20, dtype=torch.int64)) is False


def test_checkpoint_reader_round_trip(tmp_path):
    """models/checkpoint.py: a synthetic model written in the Hugging Face / AutoAWQ safetensors layout reads back
    tensor-for-tensor, config.json gives the shape back, and a model without a checkpoint refuses to allocate
    unless synthetic weights are asked for."""
    import torch
    from umbrella_b200.models import checkpoint as ck
    from umbrella_b200.models import llama as L
    from umbrella_b200.models.auto_model import AutoModelLM
    shape = L.LlamaShape(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=512, rope_scaling=None, rope_theta=10000.0)
    for awq in (True, False):
        d = str(tmp_path / ("awq" if awq else "fp16"))
        ck.write_synthetic_checkpoint(d, shape, awq, seed=5, shard_layers=2)
        got_shape, got_awq = ck.shape_from_config(d)
        assert got_awq == awq and got_shape == shape
        idx = ck.SafetensorsIndex(d)
        head, ref_head = ck.read_head(idx, shape.tie_word_embeddings), L.synth_head(shape, 5)
        assert all(torch.equal(head[k], ref_head[k]) for k in ("embed_tokens", "lm_head", "norm"))
        for i in range(shape.num_hidden_layers):
            w, ref = ck.read_layer(idx, i, awq), L.synth_layer(shape, i, awq, 5)
            ck.check_layer_shapes(shape, w, awq, i)
            for name in L.LINEAR_NAMES:
                pairs = zip(w[name], ref[name]) if awq else [(w[name], ref[name])]
                assert all(torch.equal(a, b) for a, b in pairs)
            assert torch.equal(w["input_layernorm"], ref["input_layernorm"])
    with pytest.raises(ck.CheckpointNotFound):
        AutoModelLM.from_pretrained("meta-llama/Llama-3.2-1B-Instruct", max_length=64, device="cuda:0").alloc()
    with pytest.raises(ValueError):
        ck.check_layer_shapes(L.LLAMA32_1B, ck.read_layer(ck.SafetensorsIndex(d), 0, False), False, 0)


def test_budget_tree_is_the_best_tree_of_its_size():
    """generate_budget_tree (the B200-tuned growmaps): level order, depth limit, closed under parent / elder sibling,
    and no tree of the same size and depth has a higher expected accept length (brute force over small budgets)."""
    import itertools
    from umbrella_b200.sequoia_utils import (DEFAULT_ACC, expected_accept_length, generate_budget_tree,
                                             generate_sequoia_tree, load_growmap)

    def best_by_enumeration(budget, depth):
        # all candidate nodes as rank paths; a tree = a set closed under prefix and under "smaller last rank"
        paths = [()]
        for d in range(1, depth + 1):
            paths += list(itertools.product(range(len(DEFAULT_ACC)), repeat=d))
        prob = {p: float(np.prod([DEFAULT_ACC[r] for r in p])) if p else 1.0 for p in paths}
        # the optimum of a sum of per-node values over closed sets where every node's value <= its predecessors' is the
        # top-`budget` values (ties aside); check that against the greedy result
        return sum(sorted(prob.values(), reverse=True)[:budget])

    for budget, depth in [(1, 0), (2, 1), (5, 2), (8, 2), (13, 4), (16, 4), (16, 3)]:
        gm = generate_budget_tree(budget, depth)
        assert gm["size"] == budget and max(gm["depth"]) <= depth
        for v, succ in enumerate(gm["Successors"]):
            assert all(c > v for c in succ) and succ == sorted(succ)
        assert abs(expected_accept_length(gm) - best_by_enumeration(budget, depth)) < 1e-9
    from umbrella_b200.sequoia_utils import tune_budget_tree
    shipped = load_growmap("sequoia_tree-3x4.json")
    same_depth = load_growmap("b200_sequoia_tree-16-depth4.json")
    assert same_depth["Successors"] == generate_budget_tree(16, 4)["Successors"]
    assert same_depth["size"] == 16 and max(same_depth["depth"]) == max(shipped["depth"])   # same number of draft levels
    assert expected_accept_length(same_depth) > expected_accept_length(shipped) * 1.06
    assert expected_accept_length(generate_budget_tree(13, 4)) >= expected_accept_length(generate_sequoia_tree(3, 4))
    # the packaged tuned tree is the optimum of the measured two-parameter cost model: one level fewer
    tuned, rate = tune_budget_tree(16, fixed_ms=10.0, draft_forward_ms=1.05)
    assert load_growmap("b200_sequoia_tree-16.json")["Successors"] == tuned["Successors"]
    assert tuned["size"] == 16 and max(tuned["depth"]) == 3
    shipped_rate = 1e3 * expected_accept_length(shipped) / (10.0 + 1.05 * 4)
    assert rate > 1.08 * shipped_rate
    # a free draft model wants the deepest tree, an expensive one the shallowest
    assert max(tune_budget_tree(16, 10.0, 0.0)[0]["depth"]) > max(tune_budget_tree(16, 10.0, 20.0)[0]["depth"])


# ---- acceptance-rate measurement (examples/construct_sequoia.py:58-90; sequoia_utils.measure_acceptance_rate) ---------
class _TableLM:
    """Bigram-table model with the reference's model surface: logits of a row depend on that row's token only."""
    device = "cpu"

    def __init__(self, table):
        self.table, self.cleared = table, 0

    def inference(self, input_ids, position_ids, attention_mask, storage_ids=None):
        assert attention_mask.shape == (input_ids.shape[1], input_ids.shape[1])
        return self.table[input_ids]

    def clear(self):
        self.cleared += 1


def test_measure_acceptance_rate_counts_rank_matches():
    import torch
    from umbrella_b200 import sequoia_utils as su
    g = torch.Generator().manual_seed(5)
    V, width = 50, 4
    target = torch.randn(V, V, generator=g)
    # the draft ranks the target's arg-max at rank (token % 5): ranks 0..3 are hits of that rank, 4 is a miss
    draft = torch.randn(V, V, generator=g) * 0.01
    best = target.argmax(-1)
    for t in range(V):
        order = draft[t].argsort(descending=True)
        r = t % 5
        draft[t, best[t]] = (draft[t, order[r - 1]] + draft[t, order[r]]) / 2 if r > 0 else draft[t, order[0]] + 1.0
        if r > 0 and best[t] in order[:r].tolist():
            draft[t, best[t]] = draft[t].max() + 1.0        # already ahead: settle it at rank 0 (bookkept below)
    data = [(torch.randint(0, V, (7,), generator=g), torch.randint(0, V, (9,), generator=g)) for _ in range(6)]
    tm, dm = _TableLM(target), _TableLM(draft)
    raw = su.measure_acceptance_rate(dm, tm, data, temperature=0.0, width=width, normalize=False)
    # brute force over the rows that predict a continuation token
    want = [0.0] * width
    rows = 0
    for ctx, cont in data:
        ids = torch.cat([ctx, cont])
        for i in range(len(ctx) - 1, len(ids) - 1):
            tok = int(ids[i])
            rank = draft[tok].argsort(descending=True).tolist().index(int(best[tok]))
            if rank < width:
                want[rank] += 1
            rows += 1
    assert raw == want and rows == 54
    rate = su.measure_acceptance_rate(dm, tm, data, temperature=0.0, width=width)
    assert rate == pytest.approx([w / rows for w in want]) and sum(rate) <= 1.0
    assert tm.cleared == 12 and dm.cleared == 12
    # the vector feeds the tree builder
    gm = su.generate_sequoia_tree(3, 4, acc=[max(r, 1e-6) for r in rate])
    assert gm["size"] == 13


def test_measure_acceptance_rate_stochastic_rule():
    import torch
    from umbrella_b200 import sequoia_utils as su
    g = torch.Generator().manual_seed(9)
    V = 40
    target, draft = torch.randn(V, V, generator=g) * 3, torch.randn(V, V, generator=g) * 3
    data = [(list(range(5)), list(range(5, 17)))]
    T, p = 0.6, 0.9
    rate = su.measure_acceptance_rate(_TableLM(draft), _TableLM(target), data, temperature=T, topp=p,
                                      repetition_penalty=1.0, width=V)
    # with every rank counted the kept mass is 1 per row
    assert sum(rate) == pytest.approx(1.0, abs=1e-5)
    # nucleus: the kept set is the smallest prefix of the sorted distribution whose mass reaches topp
    dist = su.target_distribution(target[:3], T, p)
    full = torch.softmax(target[:3] / T, -1)
    for r in range(3):
        sp = full[r].sort(descending=True).values
        k = int(((sp.cumsum(0) - sp) < p).sum())
        assert int((dist[r] > 0).sum()) == k and dist[r].sum().item() == pytest.approx(1.0, abs=1e-6)
    # the penalty moves mass away from tokens already in the prefix
    pen = su.measure_acceptance_rate(_TableLM(draft), _TableLM(target), data, temperature=T, topp=1.0,
                                     repetition_penalty=1.5, width=V)
    assert sum(pen) == pytest.approx(1.0, abs=1e-5) and pen != pytest.approx(rate)
    # greedy rule on the same tables = one-hot rows
    assert su.target_distribution(target[:2], 0.0).sum().item() == 2.0


# ---- streaming front-end loop (static_speculation_engine.py:470-551) ---------------------------------------------------
class _ScriptedEngine:
    """The members stream_generation touches, with a verify() that commits a scripted number of tokens per step."""
    stop_distance = 4

    def __init__(self, vocab, per_step, stop_after=None, room=10_000):
        import torch
        self.vocab, self.per_step, self.stop_after, self.room = vocab, per_step, stop_after, room
        self.tokens = torch.zeros(1, 4096, dtype=torch.int64)
        self.num_nodes, self.step, self.was_reset = 10, 0, False

    def validate_status(self):
        return self.num_nodes <= self.room

    def build_tree(self):
        pass

    def verify(self):
        k = self.per_step[self.step % len(self.per_step)]
        for i in range(k):
            self.tokens[0, self.num_nodes + i] = (self.num_nodes + i) % len(self.vocab)
        self.num_nodes += k
        self.step += 1
        return self.stop_after is None or self.step < self.stop_after

    def _decode(self, ids):
        return "".join(self.vocab[i] for i in ids)

    def reset(self):
        self.was_reset = True


def _stream_cases():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream.json")) as f:
        gold = json.load(f)
    return [(gold["vocab"], gold["start"], c) for c in gold["cases"]]


@pytest.mark.parametrize("vocab,start,case", _stream_cases())
def test_stream_generation_follows_the_reference_protocol(vocab, start, case):
    """tests/golden/stream.json: the texts the reference's own generate_stream yields (both engines, run on CPU by
    oracle/make_golden_stream.py over the same scripted engine) -- token limit, engine stop, sentence end inside the
    stop_distance window, max_new_tokens below stop_distance, validate_status ending the loop."""
    from umbrella_b200.speculation.speculation_utils import stream_generation
    assert start == 10
    eng = _ScriptedEngine(vocab, case["per_step"], case["stop_after"], case["room"])
    eng.stop_distance = case["stop_distance"]
    got = list(stream_generation(eng, case["max_new"]))
    assert [t for t, _ in got] == case["texts"]
    assert eng.num_nodes - start == case["committed"] and eng.was_reset
    assert all(p.startswith("Output Tokens ") and "| Avg Accept Tokens " in p and p.endswith(" ms ") for _, p in got)
    texts = [t for t, _ in got]
    assert all(b.startswith(a) for a, b in zip(texts, texts[1:]))          # released text only ever grows
    assert texts[-1] == eng._decode(eng.tokens[0, start:eng.num_nodes].tolist()).strip()



def test_dynamic_scripted_acceptance_picks_the_scripted_child_rank():
    """DynamicScriptedAcceptance against a per-node loop: the rank-r child by draft score replaces the sampled token."""
    import torch
    from types import SimpleNamespace
    from umbrella_b200.speculation.scripted import DynamicScriptedAcceptance
    rng = np.random.default_rng(3)
    width, depth = 4, 5
    N, n = width * depth + 1, 17
    acc = [0.5, 0.25, 0.15]
    script = DynamicScriptedAcceptance(N, seed=8, max_iters=6, device="cpu", acc=acc)
    replaced = kept = 0
    for it in range(9):                                     # wraps around max_iters
        parents = np.zeros(N, dtype=np.int32)
        score = np.zeros(N, dtype=np.float32)
        for lvl in range(depth):
            first = 1 + lvl * width
            prev = [0] if lvl == 0 else list(range(first - width, first))
            for w in range(first, first + width):
                parents[w] = rng.choice(prev)
                score[w] = score[parents[w]] - rng.random()
        tokens = torch.zeros(1, 64, dtype=torch.int64)
        tokens[0, n:n + N] = torch.from_numpy(rng.integers(1000, 2000, N))
        eng = SimpleNamespace(tree_size=N, num_nodes=n, parents=torch.from_numpy(parents),
                              tree_score=torch.from_numpy(score), tokens=tokens)
        sampled = torch.arange(N, dtype=torch.int64)        # stands for the target's own tokens (match no child)
        want = script.want[it % 6].tolist()
        got = script(eng, sampled).tolist()
        for v in range(N):
            kids = sorted([w for w in range(1, N) if parents[w] == v], key=lambda w: (-score[w], w))
            if 0 <= want[v] < len(kids):
                assert got[v] == int(tokens[0, n + kids[want[v]]]), (it, v)
                replaced += 1
            else:
                assert got[v] == v
                kept += 1
    assert replaced > 20 and kept > 20
    # the draws follow acc: rank r with probability acc[r], none with the rest
    big = DynamicScriptedAcceptance(50, seed=1, max_iters=400, device="cpu", acc=acc).want.flatten()
    freq = [float((big == r).float().mean()) for r in (0, 1, 2, -1)]
    assert freq == pytest.approx(acc + [1 - sum(acc)], abs=0.02)


def test_graph_inference_surface():
    """Model surface of the reference's examples (examples/bench.py:30-47): graph_inference == inference for plain models,
    storage-id placement for cuda_graph=True models, initialize_cuda_graph clears the cache."""
    import torch
    from types import SimpleNamespace
    from umbrella_b200.models.llama import Llama
    calls = []

    def fake(storage_semantics):
        kv = SimpleNamespace(kv_offset=40, clear=lambda: calls.append("clear"))
        m = SimpleNamespace(kv_cache=kv, storage_id_semantics=storage_semantics)
        m.inference = lambda ids, pos, mask, store: calls.append(("inference", kv.kv_offset, ids.shape[1])) or "logits"
        m.clear = kv.clear
        return m

    ids = torch.zeros(1, 3, dtype=torch.long)
    plain = fake(False)
    assert Llama.graph_inference(plain, input_ids=ids, position_ids=None, attention_mask=None,
                                 storage_ids=torch.arange(7, 10)) == "logits"
    assert calls[-1] == ("inference", 40, 3)                     # appended at kv_offset, storage values ignored
    graph = fake(True)
    Llama.graph_inference(graph, input_ids=ids, storage_ids=torch.arange(7, 10), position_ids=None, attention_mask=None)
    assert calls[-1] == ("inference", 7, 3)                      # rows written where storage_ids says
    with pytest.raises(ValueError):
        Llama.graph_inference(graph, input_ids=ids, storage_ids=torch.tensor([7, 9, 10]))
    with pytest.raises(TypeError):                               # keyword arguments only: the reference's two orders differ
        Llama.graph_inference(graph, ids, torch.arange(7, 10))
    Llama.initialize_cuda_graph(graph, [1, 13], n_warmups=2)
    assert calls[-1] == "clear" and graph.graph_lengths == [1, 13]


def test_packaged_configs_build_engines():
    """configs/*.json (the BASELINE workloads in the reference's config schema) go through AutoEngine.from_config; the
    trees they name are packaged."""
    import glob
    import json
    from umbrella_b200.models.llama import KNOWN_MODELS
    from umbrella_b200.sequoia_utils import load_growmap
    from umbrella_b200.speculation.auto_engine import AutoEngine
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "configs", "*.json")))
    assert len(files) >= 5
    for f in files:
        cfg = json.load(open(f))
        assert KNOWN_MODELS.get(cfg["model"]) and KNOWN_MODELS.get(cfg["draft_model"]), f
        eng = AutoEngine.from_config("cuda:0", **cfg)          # construction is host-only; initialize() needs the GPU
        assert type(eng).__name__ == {"static": "StaticSpeculationEngine", "dynamic": "DynamicSpeculationEngine"}[cfg["engine"]]
        assert eng.max_length == cfg["max_length"] and eng.temperature == cfg["temperature"]
        if cfg["engine"] == "static":
            assert load_growmap(cfg["growmap_path"])["size"] >= 13
        else:
            assert (eng.tree_width, eng.tree_depth, eng.num_beams) == (cfg["width"], cfg["depth"], cfg["num_beams"])


def test_kv_host_layers_config_goes_to_the_paged_target_only():
    """An engine config with "kv_host_layers" builds the target with the paged class and keeps the key away from the
    draft (whose resident class refuses it)."""
    import json
    from umbrella_b200.speculation.speculation_utils import split_model_config
    cfg = {"num_cache_layers": 0, "exit_layer": 16, "cuda_graph": True}
    assert split_model_config(cfg, target_offload=False) == (cfg, cfg, False)
    assert split_model_config(cfg, target_offload=True)[2] is True
    d, t, off = split_model_config(dict(cfg, kv_host_layers=40, kv_staging_layers=3), target_offload=False)
    assert d == cfg and t["kv_host_layers"] == 40 and t["kv_staging_layers"] == 3 and off is True
    assert split_model_config(dict(cfg, kv_host_layers=0), target_offload=False)[2] is False
    # the engines use it where they build their models
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("static_speculation_engine.py", "dynamic_speculation_engine.py"):
        src = open(os.path.join(root, "umbrella_b200", "speculation", name)).read()
        assert "split_model_config(self.config" in src and "draft_model.alloc(**draft_cfg)" in src
    from umbrella_b200.speculation.auto_engine import AutoEngine
    eng = AutoEngine.from_config("cuda:0", **json.load(open(os.path.join(root, "configs", "b200_stochastic_32k_host_kv.json"))))
    assert eng.config["kv_host_layers"] == 40


@pytest.mark.parametrize("engine", ["static", "dynamic"])
def test_speculative_decoding_prints_what_the_reference_prints(engine, monkeypatch, capsys):
    """The interactive loop of both engines (static...:211-255, dynamic...:167-213) over the scripted engine of
    tests/golden/stream.json: same words on stdout, same (dec_len, steps) as the reference's own loops."""
    import torch
    if engine == "static":
        from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine as cls
    else:
        from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine as cls
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for vocab, start, case in _stream_cases():
        eng = _ScriptedEngine(vocab, case["per_step"], case["stop_after"], case["room"])
        eng.stop_distance = case["stop_distance"]
        capsys.readouterr()
        dec_len, seconds, steps = cls.speculative_decoding(eng, max_new_tokens=case["max_new"])
        assert capsys.readouterr().out == case["printed"], case
        assert (dec_len, steps) == (case["dec_len"], case["steps"]) and seconds >= 0


@pytest.mark.parametrize("engine", ["static", "dynamic"])
def test_generate_answers_what_the_reference_answers(engine, monkeypatch):
    """One API request through ``generate`` (static...:373-434, dynamic...:343-409) over the scripted engine: the same
    generated_text / generated_tokens / avg_accept_tokens as the reference's own ``generate`` (tests/golden/stream.json);
    ``large_model_steps`` and ``decode_seconds`` are additions."""
    import types
    import torch
    if engine == "static":
        from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine as cls
    else:
        from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine as cls
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for vocab, start, case in _stream_cases():
        eng = _ScriptedEngine(vocab, case["per_step"], case["stop_after"], case["room"])
        eng.num_nodes, eng.device = 0, "cpu"
        eng.update_generation_args = lambda **kw: None

        def _prefill(input_ids, eng=eng):
            assert input_ids.shape == (1, 3) and input_ids.dtype == torch.int64
            eng.num_nodes = start
            return True
        eng._prefill = _prefill
        for name in ("_empty", "_start"):
            if hasattr(cls, name):
                setattr(eng, name, types.MethodType(getattr(cls, name), eng))
        got = cls.generate(eng, input_ids=[1, 2, 3], max_new_tokens=case["max_new"], temperature=0.0)
        want = case["generate"]
        assert {k: got[k] for k in want} == want, case
        assert got["time_per_output_token"] >= 0 and got["large_model_steps"] >= 1 and eng.was_reset
    # nothing to generate: the reference's empty answer (static...:381-396)
    eng = _ScriptedEngine(["a"], [1])
    eng.update_generation_args = lambda **kw: None
    for name in ("_empty", "_start"):
        if hasattr(cls, name):
            setattr(eng, name, types.MethodType(getattr(cls, name), eng))
    for req in (dict(input_ids=[], max_new_tokens=4), dict(context="", max_new_tokens=4), dict(input_ids=[1], max_new_tokens=0)):
        out = cls.generate(eng, **req)
        assert (out["generated_text"], out["generated_tokens"], out["avg_accept_tokens"], out["time_per_output_token"]) \
            == ("", [], 0, 0)

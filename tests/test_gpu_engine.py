"""GPU parity: the static speculative-decoding engine end to end on small synthetic models.

Losslessness KAT: greedy speculative decoding must emit exactly the tokens of plain
greedy autoregressive decoding of the target model -- checked bit-exactly against an AR loop over
the same CUDA forward, and against the CPU oracle model wherever the oracle's top-2 margin is safe.
Also replays scripted acceptance and checks accept lengths against the oracle's verify_step."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import llama as ollama
from oracle import tree as otree

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
TARGET = ollama.LlamaShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                           num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)
DRAFT = ollama.LlamaShape(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4,
                          num_key_value_heads=1, vocab_size=2048, tie_word_embeddings=True, rope_scaling=None,
                          rope_theta=10000.0)


def _engine(tree="sequoia_tree-3x4.json", S=384, draft_same=False, graph=True, lazy=True):
    from umbrella_b200.models.llama import SyntheticLlama
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    eng = StaticSpeculationEngine("draft", "target", device=DEV, growmap_path=tree, max_length=S, safe_buffer=48,
                                  graph_iteration=graph, graph_bucket=128, lazy_leaf_kv=lazy)
    target = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV)
    draft = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV) if draft_same else \
        SyntheticLlama(DRAFT, max_length=S, awq=False, seed=22, device=DEV)
    eng.initialize(draft_model=draft, target_model=target)
    eng._set_eos([])
    return eng


def _ar_greedy(model, prompt, n_new):
    """Plain greedy decoding with the same CUDA forward, one token at a time."""
    from umbrella_b200 import ops
    model.clear()
    P = prompt.numel()
    ids = prompt.to(DEV)
    logits = model.prefill_forward(ids, torch.arange(P, device=DEV), last_only=True)
    out = [int(ops.argmax(logits)[0])]
    bits = torch.ones(1, 1, dtype=torch.int32, device=DEV)
    for i in range(n_new - 1):
        tok = torch.tensor([out[-1]], device=DEV)
        pos = torch.tensor([P + i], device=DEV)
        logits = model.verify_forward(tok, pos, P + i, bits)
        out.append(int(ops.argmax(logits)[0]))
    model.clear()
    return out


@pytest.mark.parametrize("tree,draft_same,graph", [("sequoia_tree-3x4.json", False, False),
                                                   ("sequoia_tree-3x4.json", False, True),
                                                   ("8b_sequoia_tree-5x6.json", True, True),
                                                   ("b200_sequoia_tree-16.json", False, True),      # tune_budget_tree(16, 10.0, 1.05): depth 3
                                                   ("b200_sequoia_tree-16-depth4.json", True, False)])
def test_greedy_speculation_is_lossless(tree, draft_same, graph):
    """graph=True runs every iteration as one captured CUDA graph (several buckets of 128 positions)."""
    eng = _engine(tree, draft_same=draft_same, graph=graph)
    g = torch.Generator().manual_seed(3)
    prompt = torch.randint(3, TARGET.vocab_size, (40,), generator=g)
    n_new = 48
    ref = _ar_greedy(eng.target_model, prompt, n_new + 8)
    out = eng.generate(input_ids=prompt.tolist(), max_new_tokens=n_new)
    toks = out["generated_tokens"]
    assert len(toks) >= n_new
    assert toks == ref[:len(toks)]
    if draft_same:                                   # identical draft -> whole chains accepted
        assert out["avg_accept_tokens"] > 2.0
    # engine state is reset after generate(), as the reference does (static...:433)
    assert eng.num_nodes == 0 and eng.target_model.kv_cache.kv_offset == 0
    # oracle cross-check of the first tokens
    oracle = ollama.OracleLlama(TARGET, 384, awq=True, seed=21)
    P = prompt.numel()
    lg = oracle.inference(prompt[None], torch.arange(P)[None], torch.from_numpy(otree.make_causal_mask(P)))[0, -1]
    top2 = lg.topk(2).values
    if float(top2[0] - top2[1]) > 8e-3 * float(lg.abs().max()):
        assert int(lg.argmax()) == toks[0]


def test_greedy_speculation_is_lossless_at_long_context():
    """Prompt of 8250 tokens: the target's verify attention (26 query rows, head_dim 128) runs on the tensor-memory
    long-context kernel with the device-resident prefix length of the iteration graph; output must equal plain greedy
    decoding through the same forward."""
    eng = _engine(S=8704, graph=True)
    g = torch.Generator().manual_seed(17)
    prompt = torch.randint(3, TARGET.vocab_size, (8250,), generator=g)
    n_new = 24
    ref = _ar_greedy(eng.target_model, prompt, n_new + 8)
    out = eng.generate(input_ids=prompt.tolist(), max_new_tokens=n_new)
    toks = out["generated_tokens"]
    assert len(toks) >= n_new and toks == ref[:len(toks)]


def test_scripted_acceptance_matches_oracle_bookkeeping():
    """Scripted run: accept lengths / tokens / positions follow the oracle's verify_step exactly."""
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    eng = _engine(graph=False)          # step-wise inspection needs the eager build_tree / verify split
    gm = eng.growmap
    succ = gm["Successors"]
    parents = otree.successors_to_parents(succ)
    mask = np.asarray(gm["mask"], dtype=bool)
    depth = np.asarray(gm["depth"])
    script = ScriptedAcceptance(succ, seed=5, max_iters=64, device=DEV)
    eng.acceptance_script = script
    g = torch.Generator().manual_seed(9)
    prompt = torch.randint(3, TARGET.vocab_size, (1, 33), generator=g).to(DEV)
    assert eng._prefill(prompt)
    st = otree.VerifyState(eng.max_length, depth)
    total = 0
    for it in range(12):
        n = eng.num_nodes
        eng.build_tree()
        st.tokens = eng.tokens[0].cpu().numpy().copy()
        st.position_ids = eng.position_ids[0].cpu().numpy().copy()
        st.num_nodes = n
        row = script.choice[script.it % script.max_iters].cpu().numpy()
        cont = eng.verify()
        sampled = script.last.cpu().numpy()           # exactly what the accept kernel consumed
        spec = st.tokens[n:n + len(row)]
        assert np.array_equal(sampled[row >= 0], spec[row[row >= 0]])
        cont_o, path, alen = otree.verify_step(st, sampled, parents, mask, mask.sum(-1), [])
        assert alen == eng.last_accept_length and st.num_nodes == eng.num_nodes and cont == cont_o
        assert np.array_equal(eng.tokens[0, :n + alen].cpu().numpy(), st.tokens[:n + alen])
        assert np.array_equal(eng.position_ids[0, :eng.num_nodes + len(row)].cpu().numpy(),
                              st.position_ids[:st.num_nodes + len(row)])
        assert eng.target_model.kv_cache.kv_offset == eng.num_nodes == eng.draft_model.kv_cache.kv_offset
        total += alen
    assert total / 12 > 1.5      # DEFAULT_ACC on the 3x4 tree: ~3.4 expected


def test_graph_iteration_equals_eager_under_script():
    """Same scripted acceptance through the captured-graph iteration and the eager path: identical
    tokens, accept lengths and step counts."""
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    outs = []
    g = torch.Generator().manual_seed(4)
    prompt = torch.randint(3, TARGET.vocab_size, (70,), generator=g).tolist()
    for graph in (False, True):
        eng = _engine(graph=graph)
        eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=8, max_iters=256, device=DEV)
        outs.append(eng.generate(input_ids=prompt, max_new_tokens=160))
    assert outs[0]["generated_tokens"] == outs[1]["generated_tokens"]
    assert outs[0]["large_model_steps"] == outs[1]["large_model_steps"]
    assert outs[0]["avg_accept_tokens"] > 2.5


@pytest.mark.parametrize("graph", [False, True])
def test_lazy_leaf_kv_equals_reference_schedule(graph):
    """lazy_leaf_kv (no draft forward for the leaf level; an accepted leaf's K/V is recomputed by the next
    iteration's first draft step) against the reference schedule (one draft forward per level): same scripted
    acceptance -> identical tokens and step counts, and the draft cache holds the same rows (fp16 noise: the
    recomputed row sees its keys at compacted positions) except the last committed one, which is still pending."""
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    g = torch.Generator().manual_seed(12)
    prompt = torch.randint(3, TARGET.vocab_size, (1, 45), generator=g).to(DEV)
    caches, toks, steps = [], [], []
    for lazy in (False, True):
        eng = _engine(draft_same=True, graph=graph, lazy=lazy)
        eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=3, max_iters=64, device=DEV)
        assert eng._prefill(prompt)
        for _ in range(14):
            eng.build_tree()
            eng.verify()
        n = eng.num_nodes
        toks.append(eng.tokens[0, :n + 1].cpu())
        steps.append(n)
        caches.append((eng.draft_model.kv_cache.k_cache[:, :n - 1].float().cpu(),
                       eng.draft_model.kv_cache.v_cache[:, :n - 1].float().cpu()))
        assert eng.draft_model.kv_cache.kv_offset == n
    assert steps[0] == steps[1] and torch.equal(toks[0], toks[1])
    for a, b in zip(caches[0], caches[1]):
        assert torch.allclose(a, b, atol=4e-3, rtol=2e-2)


def test_lazy_leaf_kv_keeps_draft_quality():
    """Greedy decoding with the draft identical to the target: whole chains are accepted with either draft
    schedule, i.e. the lazily recomputed rows are as good as the eagerly computed ones."""
    g = torch.Generator().manual_seed(6)
    prompt = torch.randint(3, TARGET.vocab_size, (50,), generator=g).tolist()
    outs = [_engine(draft_same=True, graph=True, lazy=lazy).generate(input_ids=prompt, max_new_tokens=96)
            for lazy in (False, True)]
    assert outs[0]["generated_tokens"][:96] == outs[1]["generated_tokens"][:96]
    assert abs(outs[0]["large_model_steps"] - outs[1]["large_model_steps"]) <= 2
    assert outs[1]["avg_accept_tokens"] > 3.0


def test_append_after_lazily_skipped_leaf_row():
    """with lazy_leaf_kv the draft K/V row of an accepted last-level node is written by the NEXT
    build_tree; a multi-turn ``_append`` in between must re-derive it.  After appends that follow leaf accepts, every
    committed row of the draft cache must equal (fp16 noise) what a fresh causal pass over the same tokens gives --
    a stale row (never written, or holding another tree node's K/V) is far outside that tolerance."""
    from umbrella_b200.models.llama import SyntheticLlama
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    g = torch.Generator().manual_seed(31)
    prompt = torch.randint(3, TARGET.vocab_size, (1, 40), generator=g).to(DEV)
    extra = [torch.randint(3, TARGET.vocab_size, (1, 7), generator=g).to(DEV) for _ in range(3)]
    eng = _engine(draft_same=True, graph=False, lazy=True)
    # rank-0 children always accepted: the last-level node is accepted in every iteration
    eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=3, max_iters=64, device=DEV, acc=[1.0, 0.0, 0.0, 0.0])
    assert eng._prefill(prompt)
    for turn in range(3):
        for _ in range(4):
            eng.build_tree()
            eng.verify()
        assert eng.last_accept_length == eng.tree_depth      # a leaf was accepted: its draft row is pending
        assert eng._append(extra[turn])
    n = eng.num_nodes
    assert eng.draft_model.kv_cache.kv_offset == n == eng.target_model.kv_cache.kv_offset
    fresh = SyntheticLlama(TARGET, max_length=384, awq=True, seed=21, device=DEV)
    fresh.prefill_forward(eng.tokens[0, :n], torch.arange(n, device=DEV), want_logits=False)
    for got, ref in ((eng.draft_model.kv_cache.k_cache, fresh.kv_cache.k_cache),
                     (eng.draft_model.kv_cache.v_cache, fresh.kv_cache.v_cache)):
        assert torch.allclose(got[:, :n].float(), ref[:, :n].float(), atol=6e-3, rtol=3e-2)

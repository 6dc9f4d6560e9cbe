"""GPU parity: the dynamic (SpecExec-style) engine -- tree growth kernel, chunked tree forward, verify.

  * umb_dyn_tree_expand against the oracle's restatement of dynamic_speculation_engine.py:235-248;
  * the whole engine (build_tree + verify + gathers) replaying the runs the REFERENCE's own
    DynamicSpeculationEngine produced on CPU over bigram-table LMs (tests/golden/accept_dynamic.npz):
    tree tokens, parents, ancestor masks, committed tokens, positions, KV rows, counters -- bit-exact;
  * losslessness on small synthetic Llama models, including a tree of more than 128 nodes (the C
    forward splits it into chunks)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import llama as ollama
from oracle import tree as otree

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


@pytest.mark.parametrize("rows,beams,width,seed", [(1, 6, 4, 0), (4, 6, 4, 1), (16, 24, 16, 2), (21, 24, 21, 3),
                                                   (8, 8, 8, 4)])
def test_dyn_tree_expand_vs_oracle(rows, beams, width, seed):
    from umbrella_b200 import ops
    rng = np.random.default_rng(seed)
    N = 1 + rows + width + 40
    words = (N + 31) // 32
    level_first, n = (0 if rows == 1 else 1 + 7), 23
    logits = np.sort(rng.standard_normal((rows, beams)).astype(np.float32) * 3, axis=-1)[:, ::-1].copy()
    ids = rng.integers(0, 50000, (rows, beams)).astype(np.int64)
    hist = np.zeros(N, dtype=np.float32)
    hist[level_first:level_first + rows] = -rng.random(rows).astype(np.float32) * 4
    bits = np.zeros((N, words), dtype=np.uint32)
    for r in range(rows):                                # arbitrary ancestor rows for the level
        i = level_first + r
        bits[i, 0] = 1
        bits[i, i >> 5] |= np.uint32(1 << (i & 31))
    score, toks, prow = otree.dynamic_tree_level(logits, ids, hist[level_first:level_first + rows], width)

    t_score = torch.from_numpy(hist.copy()).to(DEV)
    t_tokens = torch.zeros(256, dtype=torch.int64, device=DEV)
    t_parents = torch.zeros(N, dtype=torch.int32, device=DEV)
    t_bits = torch.from_numpy(bits.view(np.int32).copy()).to(DEV)
    ops.dyn_tree_expand(torch.from_numpy(logits).to(DEV), torch.from_numpy(ids).to(DEV), width, level_first, n, t_score,
                        t_tokens, t_parents, t_bits)
    new = level_first + rows
    got_tok = t_tokens.cpu().numpy()
    assert (got_tok[n + new:n + new + width] == toks).all()
    assert (t_parents.cpu().numpy()[new:new + width] == prow + level_first).all()
    np.testing.assert_allclose(t_score.cpu().numpy()[new:new + width], score, rtol=0, atol=2e-6)
    got_bits = t_bits.cpu().numpy().view(np.uint32)
    for j in range(width):
        exp = bits[level_first + prow[j]].copy()
        exp[(new + j) >> 5] |= np.uint32(1 << ((new + j) & 31))
        assert (got_bits[new + j] == exp).all()
    assert got_tok[:n + new].sum() == 0 and got_tok[n + new + width:].sum() == 0


class BigramLM:
    """GPU twin of oracle/make_golden.py's BigramLM on the engine-facing model surface: logits =
    table[input_ids]; every forward appends identifiable rows (token, position) to a real KV_Cache."""

    def __init__(self, table: np.ndarray, max_length: int):
        from umbrella_b200.attn.cache import KV_Cache
        cfg = types.SimpleNamespace(hidden_size=32, num_attention_heads=4, num_key_value_heads=1, num_hidden_layers=2,
                                    vocab_size=table.shape[0], head_dim=8)
        self.config, self.device, self.eos_tokens = cfg, torch.device(DEV), []
        self.table = torch.from_numpy(table).to(DEV)
        self.kv_cache = KV_Cache(cfg, max_length=max_length, device=DEV)

    def _append(self, ids, pos):
        q, off = ids.numel(), self.kv_cache.kv_offset
        for layer in range(2):
            self.kv_cache.k_cache[layer, off:off + q, 0, 0] = ids.to(torch.float16)
            self.kv_cache.k_cache[layer, off:off + q, 0, 1] = pos.to(torch.float16)
            self.kv_cache.v_cache[layer, off:off + q, 0, 0] = -ids.to(torch.float16)
        self.kv_cache.kv_offset = off + q

    def prefill_forward(self, ids, pos, last_only=True, want_logits=True):
        self._append(ids, pos)
        if not want_logits:
            return None
        lg = self.table[ids].float()
        return lg[-1:].clone() if last_only else lg

    def verify_forward(self, ids, pos, prefix_len, tree_bits, argmax_keys=None, want_logits=True):
        self._append(ids, pos)
        lg = self.table[ids].float()
        if argmax_keys is not None:
            argmax_keys[:ids.numel()] = (~lg.argmax(-1)) & 0xFFFFFFFF
        return lg if want_logits else None

    def gather_kv_incremental(self, indices, offset):
        self.kv_cache.gather_kv_incremental(indices, offset)

    def clear(self):
        self.kv_cache.clear()


def test_dynamic_engine_replays_reference_runs():
    from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
    z = np.load(os.path.join(GOLDEN, "accept_dynamic.npz"))
    for ci in range(3):
        pre = "c%d_" % ci
        meta = z[pre + "meta"]
        width, beams, dpt, vocab, S, n_it = map(int, meta[:6])
        eos = [int(e) for e in meta[6:]]
        N = width * dpt + 1
        draft, target = BigramLM(z[pre + "draft_table"], S), BigramLM(z[pre + "target_table"], S)
        eng = DynamicSpeculationEngine("draft", "target", device=DEV, max_length=S, safe_buffer=16, width=width,
                                       num_beams=beams, depth=dpt)
        eng.initialize(draft_model=draft, target_model=target)
        eng._set_eos(eos)
        prompt = torch.from_numpy(z[pre + "prompt"])[None].to(DEV)
        assert eng._prefill(prompt)
        for it in range(n_it):
            n0 = eng.num_nodes
            assert n0 == int(z[pre + "n0"][it])
            eng.build_tree()
            assert (eng.tokens[0, n0:n0 + N].cpu().numpy() == z[pre + "spec"][it]).all(), (ci, it)
            assert (eng.parents.cpu().numpy() == z[pre + "parents"][it]).all(), (ci, it)
            bits = eng.tree_bits.cpu().numpy().view(np.uint32)
            assert (bits == otree.tree_bits_from_mask(z[pre + "tree_mask"][it].astype(bool))).all(), (ci, it)
            cont = eng.verify()
            assert (eng.last_sampled.cpu().numpy() == z[pre + "sampled"][it]).all()
            n1 = eng.num_nodes
            assert cont == bool(z[pre + "cont"][it]) and n1 == int(z[pre + "n1"][it]), (ci, it)
            assert (eng.tokens[0, :n1 + 1].cpu().numpy() == z[pre + "tokens_after"][it][:n1 + 1]).all()
            assert (eng.position_ids[0, :n1 + N].cpu().numpy() == z[pre + "pos_after"][it][:n1 + N]).all()
            for m in (draft, target):
                assert m.kv_cache.kv_offset == n1
            k = target.kv_cache.k_cache[0, :n1, 0, :2].float().cpu().numpy()
            assert (k[:, 0] * 1000 + k[:, 1] == z[pre + "k_digest"][it][:n1]).all(), (ci, it)
            kd = draft.kv_cache.k_cache[1, :n1, 0, :2].float().cpu().numpy()
            assert (kd == k).all()
            assert int(eng.parents.abs().sum()) == 0 and float(eng.tree_score.abs().sum()) == 0.0
        assert eng.stats["verify_calls"] == n_it


TARGET = ollama.LlamaShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                           num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)
DRAFT = ollama.LlamaShape(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4,
                          num_key_value_heads=1, vocab_size=2048, tie_word_embeddings=True, rope_scaling=None,
                          rope_theta=10000.0)


@pytest.mark.parametrize("width,beams,depth,draft_same", [(4, 6, 3, False), (8, 8, 5, True), (16, 24, 9, True)])
def test_dynamic_greedy_speculation_is_lossless(width, beams, depth, draft_same):
    """Greedy speculative decoding emits exactly the target's greedy continuation.  (16, 24, 9) is a
    145-node tree: the verify forward runs as two chunks of the C forward."""
    from test_gpu_engine import _ar_greedy
    from umbrella_b200.models.llama import SyntheticLlama
    from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
    S = 640
    eng = DynamicSpeculationEngine("draft", "target", device=DEV, max_length=S, safe_buffer=48, width=width,
                                   num_beams=beams, depth=depth)
    target = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV)
    draft = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV) if draft_same else \
        SyntheticLlama(DRAFT, max_length=S, awq=False, seed=22, device=DEV)
    eng.initialize(draft_model=draft, target_model=target)
    eng._set_eos([])
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(3, TARGET.vocab_size, (1, 37), generator=g)
    want = _ar_greedy(target, prompt[0], 40)
    out = eng.generate(input_ids=prompt[0].tolist(), max_new_tokens=32)
    got = out["generated_tokens"]
    assert len(got) >= 32
    assert got[:32] == want[:32]
    if draft_same:      # the draft IS the target: every level's best beam is accepted
        assert out["avg_accept_tokens"] > 1.5

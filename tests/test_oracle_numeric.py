"""Oracle numerics (attention, RoPE, RMSNorm, AWQ format/GEMM, Llama forward) pinned against the
fixtures of oracle/make_golden.py (reference Python on CPU / transformers / vLLM packer)."""
import os

import numpy as np
import pytest
import torch

from oracle import attention as A
from oracle import awq as Q
from oracle import llama as L
from oracle import tree as T


def test_attention_matches_reference_static_cache(golden_dir):
    z = np.load(os.path.join(golden_dir, "attention.npz"))
    for ci in range(3):
        pre = "c%d_" % ci
        hq, hkv, d, prefix, n = map(int, z[pre + "meta"])
        out = A.tree_attention(z[pre + "q"], z[pre + "k"], z[pre + "v"], prefix, z[pre + "tree_mask"])
        ref = z[pre + "out"]
        assert out.shape == ref.shape == (n, hq, d)
        # reference ran in fp32 (softmax fp32): agreement to fp32 rounding
        assert np.abs(out - ref).max() < 2e-5


def test_inv_freq_and_rope_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "rope.npz"))
    for name, shape in (("70b", L.LLAMA31_70B), ("1b", L.LLAMA32_1B)):
        inv, scal = L.llama3_inv_freq(shape)
        assert np.array_equal(inv.numpy(), z[name + "_inv_freq"])
        assert scal == float(z[name + "_attention_scaling"][0])
    inv, scal = L.llama3_inv_freq(L.LLAMA31_70B)
    cos, sin = L.rope_tables(inv, scal, 96)
    assert np.array_equal(cos.numpy(), z["rope_cos"]) and np.array_equal(sin.numpy(), z["rope_sin"])
    q, k = torch.from_numpy(z["rope_q"]), torch.from_numpy(z["rope_k"])
    qe, ke = L.apply_rope(q, k, cos, sin, torch.from_numpy(z["rope_pos"]))
    assert np.array_equal(qe.numpy(), z["rope_q_out"]) and np.array_equal(ke.numpy(), z["rope_k_out"])


def test_awq_pack_matches_vllm_packer(golden_dir):
    z = np.load(os.path.join(golden_dir, "awq_pack.npz"))
    for i in range(3):
        q, packed = z["q%d" % i], z["packed%d" % i]
        assert np.array_equal(Q.pack_cols_awq(q), packed)
        assert np.array_equal(Q.unpack_cols_awq(packed), q)


def test_awq_known_answers():
    # one int32 word, hand-computed: column j sits in nibble [0,4,1,5,2,6,3,7][j]
    word = np.array([[0x76543210]], dtype=np.uint32).view(np.int32)
    assert Q.unpack_cols_awq(word).tolist() == [[0, 4, 1, 5, 2, 6, 3, 7]]
    qw = Q.pack_cols_awq(np.array([[1, 2, 3, 4, 5, 6, 7, 8]] * 4))
    qz = Q.pack_cols_awq(np.array([[8] * 8]))
    sc = np.full((1, 8), 0.5, dtype=np.float16)
    w = Q.dequantize(qw, qz, sc, group=4)
    assert w.tolist() == [[-3.5, -3.0, -2.5, -2.0, -1.5, -1.0, -0.5, 0.0]] * 4
    x = np.ones((2, 4), dtype=np.float16)
    assert Q.gemm_exact(x, qw, qz, sc, 4).tolist() == [[-14.0, -12.0, -10.0, -8.0, -6.0, -4.0, -2.0, 0.0]] * 2


def test_awq_gemm_variants_agree_within_fp16_noise():
    rng = np.random.default_rng(0)
    k, n, g, m = 512, 64, 128, 13
    qw, qz, sc = Q.random_awq(k, n, g, rng)
    x = rng.standard_normal((m, k)).astype(np.float16)
    exact = Q.gemm_exact(x, qw, qz, sc, g)
    ours = Q.gemm_fp16_weights(x, qw, qz, sc, g).astype(np.float64)
    ref = Q.gemm_reference_emulated(x, qw, qz, sc, g).astype(np.float64)
    scale = np.abs(exact).max()
    assert np.abs(ours - exact).max() / scale < 1e-3
    assert np.abs(ref - exact).max() / scale < 3e-3
    # fp16-rounded weights are within half an fp16 ulp of the exact product
    w64 = Q.dequantize(qw, qz, sc, g, np.float64)
    w16 = Q.dequantize(qw, qz, sc, g, np.float16).astype(np.float64)
    assert np.all(np.abs(w16 - w64) <= np.abs(w64) * 2.0 ** -11 + 1e-12)


def test_rmsnorm_definition():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 64, generator=g).half()
    w = (1 + 0.1 * torch.randn(64, generator=g)).half()
    y = L.rmsnorm(x, w, 1e-5)
    x64 = x.double()
    ref = x64 / torch.sqrt((x64 ** 2).mean(-1, keepdim=True) + 1e-5) * w.double()
    assert (y.double() - ref).abs().max() < 2e-3 and y.dtype == torch.float16


TINY = L.LlamaShape(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                    num_key_value_heads=2, vocab_size=320, group_size=128)


def _greedy_ar(model, prompt, steps):
    """Plain autoregressive greedy decode through OracleLlama.inference."""
    model.clear()
    S = model.max_length
    causal = torch.from_numpy(T.make_causal_mask(S))
    p = prompt.shape[1]
    logits = model.inference(prompt, torch.arange(p)[None], causal[:p], torch.arange(p))
    toks = [int(logits[0, -1].argmax())]
    for i in range(steps - 1):
        pos = p + i
        logits = model.inference(torch.tensor([[toks[-1]]]), torch.tensor([[pos]]), causal[pos:pos + 1], torch.arange(1))
        toks.append(int(logits[0, -1].argmax()))
    return toks


def test_tree_verify_is_lossless_on_oracle_model(golden_dir):
    """KAT for the whole path: verify over a Sequoia tree whose nodes hold arbitrary tokens accepts
    exactly the greedy continuation (SURVEY.md 8c 'losslessness')."""
    import json
    succ = json.load(open(os.path.join(golden_dir, "trees.json")))["shipped"]["sequoia_tree-3x4.json"]["Successors"]
    parents = T.successors_to_parents(succ)
    tmask = np.array(T.successors_to_mask(succ)).astype(bool)
    nip = tmask.sum(-1)
    depth = nip - 1
    N, S = len(succ), 64
    model = L.OracleLlama(TINY, max_length=S, awq=True, seed=5)
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(3, TINY.vocab_size, (1, 7), generator=g)
    want = _greedy_ar(model, prompt, 12)

    model.clear()
    causal = torch.from_numpy(T.make_causal_mask(S))
    p = prompt.shape[1]
    logits = model.inference(prompt, torch.arange(p)[None], causal[:p], torch.arange(p))
    st = T.VerifyState(S, depth)
    st.prefill(prompt[0].numpy(), int(logits[0, -1].argmax()))
    produced = [int(st.tokens[st.num_nodes])]
    rng = np.random.default_rng(0)
    steps = 0
    while len(produced) < 12 and steps < 12:
        n = st.num_nodes
        # "draft": put the true continuation on the first-child chain for a random number of levels,
        # random tokens elsewhere
        spec = rng.integers(3, TINY.vocab_size, N)
        spec[0] = st.tokens[n]
        lvl = int(rng.integers(0, 5))
        chain, v = [0], 0
        while succ[v] and len(chain) <= lvl:
            v = succ[v][0]; chain.append(v)
        for dd, node in enumerate(chain[1:], start=1):
            idx = len(produced) - 1 + dd
            if idx < len(want):
                spec[node] = want[idx]
        st.tokens[n:n + N] = spec
        mask = torch.from_numpy(T.dense_rows_from_compact(n, tmask, n + N))
        logits = model.inference(torch.from_numpy(st.tokens[None, n:n + N]), torch.from_numpy(st.position_ids[None, n:n + N]),
                                 mask, torch.arange(N))
        sampled = logits[0].argmax(-1).numpy()
        cont, path_abs, alen = T.verify_step(st, sampled, parents, tmask, nip, eos_tokens=[-1])
        model.gather_kv_incremental(path_abs, n)
        produced = st.tokens[p:st.num_nodes + 1].tolist()
        steps += 1
    k = min(len(produced), len(want))
    assert k >= 8 and produced[:k] == want[:k]


def test_awq_oracle_matches_upstream_kernel(golden_dir):
    """Pins the AWQ arithmetic: tests/golden/awq_upstream.npz holds outputs of the UPSTREAM kernels
    (vLLM 0.22's build of the autoawq-kernels sources the reference calls; oracle/make_golden_gpu.py, run on a
    B200).  dequantize_weights: bit-exact.  gemm_forward (split_k_iters = 8, fp16 partial sums): the oracle's
    emulation of its rounding and the exact-accumulation form both sit within fp16 noise of it."""
    path = os.path.join(golden_dir, "awq_upstream.npz")
    if not os.path.exists(path):
        pytest.skip("awq_upstream.npz not generated yet (needs a GPU box)")
    z = np.load(path)
    ci = 0
    while "c%d_meta" % ci in z:
        pre = "c%d_" % ci
        m, k, n, g = map(int, z[pre + "meta"])
        qw, qz, sc, x = z[pre + "qweight"], z[pre + "qzeros"], z[pre + "scales"], z[pre + "x"]
        deq = Q.dequantize(qw, qz, sc, g, np.float16)
        assert deq.dtype == np.float16 and (deq.view(np.uint16) == z[pre + "dequant"].view(np.uint16)).all(), ci
        y_up = z[pre + "y"].astype(np.float64)
        exact = Q.gemm_exact(x, qw, qz, sc, g)
        scale = np.abs(exact).max()
        err_up = np.abs(y_up - exact).max() / scale                   # how far the upstream kernel is from the truth
        err_ours = np.abs(Q.gemm_fp16_weights(x, qw, qz, sc, g).astype(np.float64) - exact).max() / scale
        assert err_up < 4e-3 and err_ours <= max(err_up, 1e-3), (ci, err_up, err_ours)
        emu = Q.gemm_reference_emulated(x, qw, qz, sc, g).astype(np.float64)
        assert np.abs(emu - y_up).max() / scale < 4e-3, ci
        ci += 1
    assert ci >= 4

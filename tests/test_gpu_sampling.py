"""GPU parity: the fused target sampler (umb_sample_f32) against the oracle (oracle/sampling.py: the reference's
static / dynamic stochastic verify restated in float64), bit-equal sampled ids on seeded logits wherever the
oracle's decision margin exceeds the fp32 noise of the device arithmetic; the algorithm itself is pinned against
flashinfer 0.6.11's own samplers (distribution level -- that release has no uniform_samples argument)."""
import numpy as np
import pytest
import torch

from oracle import sampling as osamp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MARGIN = 2e-5          # |running sum - u| below this may flip between float64 (oracle) and fp32 (device)


def _logits(rows, vocab, seed, fp16_grid=False):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, vocab)) * 3).astype(np.float32)
    if fp16_grid:                      # the lm_head writes float(fp16): many exact ties
        x = x.astype(np.float16).astype(np.float32)
    return x, rng


@pytest.mark.parametrize("rows,vocab,k,p,T,pen,grid", [
    (13, 128256, 32, 0.9, 0.6, 1.0, True), (13, 128256, 32, 0.9, 0.6, 1.05, True), (31, 2048, 8, 0.9, 0.7, 1.1, False),
    (64, 32000, 1, 0.5, 1.0, 1.0, False), (64, 32000, 1024, 0.95, 1.3, 1.0, True), (505, 4096, 32, 0.9, 0.6, 1.2, True),
    (5, 1000, 5000, 0.3, 0.05, 1.0, False)])
@pytest.mark.parametrize("mode", ["static", "dynamic"])
def test_sampler_vs_oracle(mode, rows, vocab, k, p, T, pen, grid):
    from umbrella_b200 import ops
    if k > ops.SAMPLE_MAX_TOPK:
        k = min(k, ops.SAMPLE_MAX_TOPK)
    x, rng = _logits(rows, vocab, rows + vocab + k, grid)
    ids = rng.integers(0, vocab, 700).astype(np.int64)
    u = rng.random((3, rows)).astype(np.float32)
    xd, ud, idd = torch.from_numpy(x).to(DEV), torch.from_numpy(u).to(DEV), torch.from_numpy(ids).to(DEV)
    kw = dict(temperature=T, top_k=k, top_p=p, penalty_ids=idd, penalty=pen, penalty_count=600)
    if mode == "static":
        ref, margin = osamp.sample_static(x, u, T, k, p, ids[:600], pen)
        out = ops.sample(xd, ops.UMB_SAMPLE_STATIC, uniforms=ud, **kw)
    else:
        ref, margin = osamp.sample_dynamic(x, u[0], T, k, p, ids[:600], pen)
        out = ops.sample(xd, ops.UMB_SAMPLE_DYNAMIC, uniforms=ud[0].contiguous(), **kw)
    out = out.cpu().numpy()
    safe = margin > MARGIN
    assert safe.mean() > 0.85          # many near-zero-width tail tokens (k = 1024, high T) sit inside the margin
    assert np.array_equal(out[safe], ref[safe]), (out[safe] != ref[safe]).sum()
    # a penalised logit must have been penalised exactly once however often its id was seen (scatter semantics)
    if pen > 1.01:
        am = ops.sample(xd, ops.UMB_SAMPLE_ARGMAX, penalty_ids=idd, penalty=pen, penalty_count=600).cpu().numpy()
        xp = osamp.apply_repetition_penalty(ids[:600], x, pen)
        top2 = np.sort(xp, axis=-1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-5
        assert np.array_equal(am[clear], xp.argmax(-1)[clear])


def test_argmax_first_occurrence_on_ties():
    from umbrella_b200 import ops
    x = np.zeros((4, 5000), dtype=np.float32)
    x[0, [17, 900]] = 3.0
    x[1, :] = -1.0
    x[2, 4999] = 1.0
    x[3, [0, 4999]] = 2.0
    out = ops.sample(torch.from_numpy(x).to(DEV), ops.UMB_SAMPLE_ARGMAX).cpu().numpy()
    assert out.tolist() == [17, 0, 4999, 0]


def test_distribution_pinned_against_flashinfer():
    """flashinfer 0.6.11's top_k_top_p_sampling_from_logits (its own Philox stream) and the fused kernel (fresh
    uniforms per draw) must both reproduce the oracle's exact distribution: the renormalised top-p nucleus of
    softmax(top-k(logits) / T); top_p_renorm_probs must keep exactly the oracle's nucleus."""
    import refkernels
    if refkernels.available() is not None:
        pytest.skip("flashinfer unavailable: " + str(refkernels.available()))
    import flashinfer
    from umbrella_b200 import ops
    x, rng = _logits(1, 128256, 7, True)
    T, k, p, M = 0.6, 32, 0.9, 20000
    exact = osamp.exact_distribution(x[0], T, k, p)
    xd = torch.from_numpy(x).to(DEV).expand(M, -1).contiguous()

    def tv(samples):
        cnt = np.bincount(samples, minlength=x.shape[1]) / len(samples)
        return 0.5 * np.abs(cnt - exact).sum()

    fi = flashinfer.sampling.top_k_top_p_sampling_from_logits(xd / T, k, p).cpu().numpy()
    u = torch.rand(3, M, device=DEV)
    ours_s = ops.sample(xd, ops.UMB_SAMPLE_STATIC, temperature=T, top_k=k, top_p=p, uniforms=u).cpu().numpy()
    ours_d = ops.sample(xd, ops.UMB_SAMPLE_DYNAMIC, temperature=T, top_k=k, top_p=p, uniforms=u[0].contiguous()).cpu().numpy()
    tvs = {"flashinfer": tv(fi), "ours_static": tv(ours_s), "ours_dynamic": tv(ours_d)}
    print("\nsampling TV distance to the oracle distribution:", tvs, "support", int((exact > 0).sum()))
    # 3 bounded rejection rounds leave the static path a little mass outside the nucleus (so does the reference)
    assert tvs["flashinfer"] < 0.03 and tvs["ours_dynamic"] < 0.03 and tvs["ours_static"] < 0.05
    probs = torch.softmax(torch.from_numpy(osamp.top_k_mask(x / T, k)).float().to(DEV), -1)
    ren = flashinfer.sampling.top_p_renorm_probs(probs, p).cpu().numpy()[0]
    keep = osamp.nucleus(osamp.softmax(osamp.top_k_mask(x.astype(np.float64) / T, k))[0], p)
    assert np.array_equal(ren > 0, keep)


@pytest.mark.parametrize("engine_kind", ["static", "dynamic"])
def test_engine_stochastic_verify_is_the_oracle_algorithm(engine_kind):
    """temperature > 0 and repetition_penalty > 1 through verify(): every sampled id of every verify equals the
    oracle's (static: the engine's fixed uniform_samples, static...:131,310; dynamic: the uniforms of that call),
    the bookkeeping stays consistent, and a seeded run is reproducible."""
    from test_gpu_engine import TARGET
    from umbrella_b200.models.llama import SyntheticLlama
    from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    S = 384
    outs = []
    for rep in range(2):
        kw = dict(device=DEV, max_length=S, safe_buffer=48, temperature=0.7, topp=0.9, topk=8, repetition_penalty=1.1)
        if engine_kind == "static":
            eng = StaticSpeculationEngine("draft", "target", growmap_path="sequoia_tree-3x4.json", **kw)
        else:
            eng = DynamicSpeculationEngine("draft", "target", width=4, num_beams=6, depth=3, **kw)
        target = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV)
        draft = SyntheticLlama(TARGET, max_length=S, awq=True, seed=21, device=DEV)
        torch.manual_seed(77)
        eng.initialize(draft_model=draft, target_model=target)
        eng._set_eos([])
        g = torch.Generator().manual_seed(2)
        prompt = torch.randint(3, TARGET.vocab_size, (1, 30), generator=g).to(DEV)
        assert eng._prefill(prompt)
        seen = []
        inner = eng._sample_target

        def spy(logits, keys, _inner=inner, _eng=eng):
            out = _inner(logits, keys)
            u = _eng.uniform_samples if engine_kind == "static" else _eng.last_uniforms
            seen.append((logits.cpu().numpy(), _eng.tokens[0, :_eng.num_nodes + 1].cpu().numpy(), u.cpu().numpy(),
                         out.cpu().numpy()))
            return out
        eng._sample_target = spy
        for _ in range(10):
            n = eng.num_nodes
            eng.build_tree()
            assert eng.verify()
            assert eng.num_nodes == n + eng.last_accept_length and eng.last_accept_length >= 1
            assert eng.target_model.kv_cache.kv_offset == eng.num_nodes == eng.draft_model.kv_cache.kv_offset
        checked = 0
        for lg, toks, u, got in seen:
            fn = osamp.sample_static if engine_kind == "static" else osamp.sample_dynamic
            ref, margin = fn(lg, u, 0.7, 8, 0.9, toks, 1.1)
            safe = margin > MARGIN
            assert np.array_equal(got[safe], ref[safe])
            checked += int(safe.sum())
        assert checked > 0.9 * sum(len(s[3]) for s in seen)
        outs.append(eng.tokens[0, :eng.num_nodes + 1].cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("rows,vocab,k", [(1, 128256, 4), (13, 128256, 24), (64, 32000, 64), (3, 1000, 1)])
def test_topk_kernel_equals_torch(rows, vocab, k):
    """umb_topk_f32 against torch.topk: same values; same ids wherever the value is unique in the row (equal values:
    lower token id first, torch leaves that order unspecified)."""
    from umbrella_b200 import ops
    x, _ = _logits(rows, vocab, 11 * rows + k, fp16_grid=(rows % 2 == 1))
    xd = torch.from_numpy(x).to(DEV)
    vals, ids = ops.topk(xd, k)
    tv, ti = xd.topk(k, dim=-1)
    assert torch.equal(vals, tv)
    assert torch.equal(xd.gather(1, ids), vals)                       # every id points at its value
    uniq = torch.ones_like(tv, dtype=torch.bool)
    uniq[:, 1:] &= tv[:, 1:] != tv[:, :-1]
    uniq[:, :-1] &= tv[:, :-1] != tv[:, 1:]
    assert torch.equal(ids[uniq], ti[uniq])
    srt = ids.clone()
    for r in range(rows):                                            # among equals: ascending token id
        for j in range(1, k):
            if vals[r, j] == vals[r, j - 1]:
                assert int(srt[r, j]) > int(srt[r, j - 1])
    toks = torch.zeros(256, dtype=torch.int64, device=DEV)
    gather = torch.randint(0, rows * k, (min(40, rows * k),), device=DEV)
    ops.store_tokens_gather(toks, ids, gather, 7, num_nodes=100)
    assert torch.equal(toks[107:107 + gather.numel()], ids.flatten()[gather])

"""GPU parity through the reference's own class surface: ``AwqLinear`` (umbrella/quantization/
awq_utils.py:5-86) and ``KV_Cache`` (umbrella/attn/cache.py:5-96) used the way LlamaAwq.layer_compute uses them,
checked against the oracle -- including the edge cases the seams imply (bias, > 128 rows, paging through
``copy``, empty accept path, append semantics that ignore the values of storage_ids)."""
import types

import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import awq as oawq
from oracle import tree as otree

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _wq_module(K, N, seed, bias=False, group=128, w_bit=4):
    """A stand-in for awq.modules.linear.WQLinear_GEMM: the attributes AwqLinear.init_parameters reads."""
    rng = np.random.default_rng(seed)
    qw, qz, sc = oawq.random_awq(K, N, 128, rng, scale=0.01)
    m = types.SimpleNamespace(in_features=K, out_features=N, w_bit=w_bit, group_size=group,
                              qweight=torch.from_numpy(qw), qzeros=torch.from_numpy(qz), scales=torch.from_numpy(sc),
                              bias=torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float16)) if bias else None)
    return m, (qw, qz, sc)


@pytest.mark.parametrize("rows,bias", [(1, False), (13, True), (200, False)])
def test_awq_linear_apply(rows, bias):
    from umbrella_b200.quantization.awq_utils import AwqLinear
    K, N = 512, 384
    mod, (qw, qz, sc) = _wq_module(K, N, rows, bias)
    lin = AwqLinear()
    lin.init_parameters(mod)                      # pinned host copies (awq_utils.py:24-26)
    assert lin.qweight.is_pinned() and not lin.qweight.is_cuda
    with pytest.raises(RuntimeError):
        lin.apply(torch.zeros(1, 1, K, dtype=torch.float16, device=DEV))      # weights not on the device yet
    lin.to(DEV)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((1, rows, K)) * 0.5).astype(np.float16)
    y = lin.apply(torch.from_numpy(x).to(DEV))
    assert y.shape == (1, rows, N) and y.dtype == torch.float16
    ref = oawq.gemm_fp16_weights(x[0], qw, qz, sc, 128).astype(np.float64)
    if bias:
        ref = (ref.astype(np.float16) + mod.bias.numpy()).astype(np.float64)
    err = np.abs(y[0].float().cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-3, err


def test_awq_linear_paging_copy_and_errors():
    """The offload pattern of LlamaAwqLayer.copy (llama_layer.py:251-258): a device-side buffer layer receives the
    weights of different host layers through ``copy`` and must compute with the new weights each time."""
    from umbrella_b200.quantization.awq_utils import AwqLinear
    K, N = 256, 256
    mods = [_wq_module(K, N, s) for s in (1, 2)]
    host = []
    for mod, _ in mods:
        h = AwqLinear()
        h.init_parameters(mod)
        host.append(h)
    buf = AwqLinear()
    buf.empty_like(mods[0][0])
    buf.to(DEV)
    x = (np.random.default_rng(0).standard_normal((1, 7, K)) * 0.5).astype(np.float16)
    xd = torch.from_numpy(x).to(DEV)
    for h, (_, (qw, qz, sc)) in zip(host, mods):
        buf.copy(h)
        y = buf.apply(xd)[0].float().cpu().numpy()
        ref = oawq.gemm_fp16_weights(x[0], qw, qz, sc, 128).astype(np.float64)
        assert np.abs(y - ref).max() / np.abs(ref).max() < 2e-3
    bad, _ = _wq_module(K, N, 3, group=64)
    with pytest.raises(ValueError):
        AwqLinear().init_parameters(bad)


def test_kv_cache_class_seams():
    from umbrella_b200.attn.cache import KV_Cache
    cfg = types.SimpleNamespace(hidden_size=512, num_attention_heads=8, num_key_value_heads=2, num_hidden_layers=3,
                                head_dim=64)
    S, P, N = 128, 37, 13
    cache = KV_Cache(cfg, max_length=S, device=DEV, zero_tail=True)
    rng = np.random.default_rng(2)
    k_all = rng.standard_normal((3, P + N, 2, 64)).astype(np.float16)
    v_all = rng.standard_normal((3, P + N, 2, 64)).astype(np.float16)
    q = rng.standard_normal((1, N, 8, 64)).astype(np.float16)
    tm = np.asarray(otree.generate_sequoia_tree(3, 4)["mask"], dtype=bool)
    mask = otree.dense_rows_from_compact(P, tm, S)
    # prefix rows, then the tree rows through compute_attention (append at kv_offset; storage_ids VALUES ignored)
    for layer in range(3):
        cache.update_kv_cache(torch.from_numpy(k_all[layer, :P]).to(DEV), torch.from_numpy(v_all[layer, :P]).to(DEV),
                              layer, torch.arange(P, device=DEV))
    assert cache.kv_offset == P
    outs = []
    for layer in range(3):
        o = cache.compute_attention(torch.from_numpy(q).to(DEV), torch.from_numpy(k_all[layer, P:])[None].to(DEV),
                                    torch.from_numpy(v_all[layer, P:])[None].to(DEV), layer,
                                    torch.full((N,), 999, device=DEV), torch.from_numpy(mask).to(DEV))
        outs.append(o.float().cpu().numpy())
    assert cache.kv_offset == P + N
    for layer in range(3):
        ref = oattn.masked_gqa_attention(q[0], k_all[layer], v_all[layer], mask[:, :P + N])
        assert np.abs(outs[layer] - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())
    # accept path [0, 2, 6]: overlapping source / destination rows, then the tail is zeroed
    want_k = cache.k_cache.cpu().numpy().copy()
    idx = np.array([P + 0, P + 2, P + 6])
    new_off = otree.kv_gather(want_k, idx, P)
    cache.gather_kv_incremental(torch.from_numpy(idx).to(DEV), P)
    assert cache.kv_offset == new_off == P + 3
    assert np.array_equal(cache.k_cache.cpu().numpy(), want_k)
    # empty accept path (EOS at the root)
    cache.gather_kv_incremental(torch.zeros(0, dtype=torch.int64, device=DEV), cache.kv_offset)
    assert cache.kv_offset == P + 3
    cache.clear()
    assert cache.kv_offset == 0 and float(cache.k_cache.abs().sum()) == 0.0


@pytest.mark.parametrize("awq", [True, False])
def test_checkpoint_loader_equals_direct_weights(tmp_path, awq):
    """safetensors on disk -> AutoModelLM.from_pretrained(<dir>).alloc() -> the same logits, bit for bit, as the
    model fed the same tensors directly (umb_awq_repack is deterministic)."""
    from umbrella_b200.models import checkpoint as ck
    from umbrella_b200.models import llama as L
    from umbrella_b200.models.auto_model import AutoModelLM
    shape = L.LlamaShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)
    d = ck.write_synthetic_checkpoint(str(tmp_path / "m"), shape, awq, seed=9)
    loaded = AutoModelLM.from_pretrained(d, max_length=128, device=DEV)
    loaded.alloc()
    assert loaded.weights_source == d and loaded.awq == awq
    direct = L.SyntheticLlama(shape, max_length=128, awq=awq, seed=9, device=DEV)
    ids = torch.randint(3, 2048, (37,), device=DEV)
    a = loaded.prefill_forward(ids, torch.arange(37, device=DEV), last_only=False)
    b = direct.prefill_forward(ids, torch.arange(37, device=DEV), last_only=False)
    assert torch.equal(a, b)

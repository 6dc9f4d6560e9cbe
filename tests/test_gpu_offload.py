"""GPU parity: the paged (host-offloaded) models produce bit-identical logits and KV rows to the resident
models over the same weights -- prefill (several 128-row chunks), tree verify, repeated forwards (the
staging ring wraps around), for every split of cached / staged layers."""
import numpy as np
import pytest
import torch

from oracle import llama as ollama
from oracle import tree as otree

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPE = ollama.LlamaShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=6, num_attention_heads=4,
                          num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)


def _run(model, prompt, tree_bits, N):
    from umbrella_b200 import ops
    outs = []
    P = prompt.numel()
    lg = model.prefill_forward(prompt, torch.arange(P, device=DEV), last_only=False)
    outs.append(lg)
    n = P
    keys = torch.zeros(N, dtype=torch.int64, device=DEV)
    for it in range(3):
        ids = torch.randint(3, SHAPE.vocab_size, (N,), generator=torch.Generator().manual_seed(it)).to(DEV)
        pos = n + torch.arange(N, device=DEV)
        lg = model.verify_forward(ids, pos, n, tree_bits, argmax_keys=keys)
        outs.append(lg)
        outs.append(ops.argmax_keys_to_ids(keys).float())
        model.kv_cache.kv_offset = n + 2            # keep two rows, as an accept of length 2 would
        n += 2
    outs.append(model.kv_cache.k_cache[:, :n].float())
    return outs


@pytest.mark.parametrize("awq,cache_layers,staging", [(True, 0, 2), (True, 2, 2), (True, 1, 3), (False, 0, 2),
                                                      (True, 4, 2), (True, 6, 2)])
def test_paged_equals_resident(awq, cache_layers, staging):
    from umbrella_b200.models.llama import SyntheticLlama
    S = 512
    tree = otree.generate_sequoia_tree(3, 4)
    tm = np.asarray(tree["mask"], dtype=bool)
    bits = torch.from_numpy(otree.tree_bits_from_mask(tm).view(np.int32)).to(DEV)
    N = tm.shape[0]
    prompt = torch.randint(3, SHAPE.vocab_size, (300,), generator=torch.Generator().manual_seed(9)).to(DEV)
    resident = SyntheticLlama(SHAPE, max_length=S, awq=awq, seed=31, device=DEV)
    want = _run(resident, prompt, bits, N)
    del resident
    paged = SyntheticLlama(SHAPE, max_length=S, awq=awq, seed=31, device=DEV,
                           offload={"num_cache_layers": cache_layers, "num_staging_layers": staging})
    got = _run(paged, prompt, bits, N)
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    n_off = SHAPE.num_hidden_layers - cache_layers
    if n_off > staging and n_off % staging == 0:
        # 4 forwards, each streams every offloaded layer exactly once (+ the initial fill of the ring)
        assert paged.copy_stats["copies"] == 4 * n_off + staging
    elif n_off > staging:
        assert paged.copy_stats["copies"] >= 4 * n_off + staging
    else:
        assert paged.copy_stats["copies"] == n_off
    torch.cuda.synchronize()


def test_auto_model_keeps_fitting_models_resident():
    from umbrella_b200.models.auto_model import AutoModelLM
    m = AutoModelLM.from_pretrained("meta-llama/Llama-3.2-1B-Instruct", offload=True, max_length=256, device=DEV)
    m.alloc(num_cache_layers=0, synthetic=True)
    assert m.num_cache_layers == m.num_layers and not m._stage
    m2 = AutoModelLM.from_pretrained("meta-llama/Llama-3.2-1B-Instruct", offload=True, max_length=256, device=DEV)
    m2.alloc(num_cache_layers=4, force_offload=True, synthetic=True)
    assert m2.num_cache_layers == 4 and len(m2._stage) == 2
    ids = torch.randint(3, 30000, (40,), device=DEV)
    a = m.prefill_forward(ids, torch.arange(40, device=DEV))
    b = m2.prefill_forward(ids, torch.arange(40, device=DEV))
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        AutoModelLM.from_pretrained("no/such-model")


def test_kv_host_pages_equal_resident_cache():
    """Host-resident KV pages (models/kv_paging.py): the K/V planes of the upper layers live in pinned host memory and
    stream through two staging planes.  Same kernels, same bytes -> prefill logits, every verify and the committed
    tokens are bit-identical to the fully resident model, and after the run the host planes hold exactly the rows the
    resident cache holds (accept-path compaction included)."""
    from oracle import llama as ollama
    from umbrella_b200.models.llama import SyntheticLlama
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    shape = ollama.LlamaShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=5, num_attention_heads=4,
                              num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)
    S = 512
    g = torch.Generator().manual_seed(3)
    prompt = torch.randint(3, shape.vocab_size, (1, 150), generator=g).to(DEV)      # > 64 rows: chunked prefill
    runs = []
    for host_layers in (0, 3):
        eng = StaticSpeculationEngine("draft", "target", device=DEV, growmap_path="sequoia_tree-3x4.json", max_length=S,
                                      safe_buffer=48, graph_iteration=False)
        target = SyntheticLlama(shape, max_length=S, awq=True, seed=5, device=DEV, kv_host_layers=host_layers)
        draft = SyntheticLlama(shape, max_length=S, awq=True, seed=5, device=DEV)
        eng.initialize(draft_model=draft, target_model=target)
        eng._set_eos([])
        eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=4, max_iters=64, device=DEV)
        assert eng._prefill(prompt)
        for _ in range(12):
            eng.build_tree()
            eng.verify()
        n = eng.num_nodes
        kc = target.kv_cache
        if host_layers:
            kc.apply_pending_gather()
            assert kc.k_cache.shape[0] == shape.num_hidden_layers - host_layers + 2
            k = torch.cat([kc.k_cache[:kc.resident_layers, :n].cpu(), kc.k_host[:, :n]], 0)
            v = torch.cat([kc.v_cache[:kc.resident_layers, :n].cpu(), kc.v_host[:, :n]], 0)
            stats = target.kv_paging_stats()
            assert stats["kv_host_layers"] == 3 and stats["h2d_bytes_per_forward"] > 0
        else:
            k, v = kc.k_cache[:, :n].cpu(), kc.v_cache[:, :n].cpu()
        runs.append((eng.tokens[0, :n + 1].cpu(), k, v))
    assert torch.equal(runs[0][0], runs[1][0])
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])

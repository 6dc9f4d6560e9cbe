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

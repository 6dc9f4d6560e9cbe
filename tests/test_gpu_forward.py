"""GPU parity: the native Llama verify forward (umb_llama_forward) on small models against
  * the oracle (oracle/llama.py: the reference's op order, fp32 maths, the reference's fp16 rounding points), run on
    the device through the same streaming helper as the 70B-shape test, and
  * the reference's own kernels (tests/refkernels.py).
north_star tolerance: target logits within 1e-3 -- read relative to max|logit| (one fp16 ulp of the top logit is
already 4.9e-4 .. 9.8e-4 of it) -- or, where the reference's own kernels sit farther than that from the same oracle on the
same inputs, their measured distance; arg-max equal wherever the oracle's top-2 margin exceeds twice the bound."""
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
TOL = 1e-3

SMALL = dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
             num_key_value_heads=2, vocab_size=2048, rope_scaling=None, rope_theta=10000.0)
SMALL64 = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
               num_key_value_heads=1, vocab_size=1024, tie_word_embeddings=True, rope_scaling=None, rope_theta=10000.0)


def _tree(name="sequoia_tree-3x4.json"):
    trees = json.load(open(os.path.join(GOLDEN, "trees.json")))["shipped"]
    return otree.growmap_from_successors(trees[name]["Successors"])


def _rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max())


@pytest.mark.parametrize("cfg,awq", [(SMALL, True), (SMALL, False), (SMALL64, False)])
def test_prefill_then_tree_verify(cfg, awq):
    import refkernels
    from umbrella_b200.models import llama as L
    shape, oshape = L.LlamaShape(**cfg), ollama.LlamaShape(**cfg)
    S, P, seed = 256, 150, 77                                   # P > 128: exercises chunked prefill
    gm = _tree()
    tm = np.asarray(gm["mask"], dtype=bool)
    N, depth = gm["size"], np.asarray(gm["depth"])
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(3, shape.vocab_size, (1, P), generator=g).to(DEV)
    tree_tokens = torch.randint(3, shape.vocab_size, (1, N), generator=g).to(DEV)
    pos_p = torch.arange(P, device=DEV)[None]
    pos_t = (P + torch.from_numpy(depth)).to(DEV)[None]
    causal = torch.from_numpy(otree.make_causal_mask(P)).to(DEV)
    dense = torch.from_numpy(otree.dense_rows_from_compact(P, tm, P + N)).to(DEV)

    model = (L.LlamaAwq if awq else L.Llama)(shape, max_length=S, device=DEV)
    model.alloc()
    res = refkernels.three_way_logits(L, ollama, shape, oshape, awq, seed, shape.num_hidden_layers, prompt, tree_tokens, pos_p,
                                      pos_t, causal, dense, model.cos_cache, model.sin_cache, DEV,
                                      on_layer=model.load_awq_layer if awq else model.load_fp16_layer, max_length=S)
    head = res["head"]
    model.load_head(head["embed_tokens"], head["lm_head"], head["norm"])
    oracle, ref = res["oracle"], res["reference"]
    model.inference(prompt, pos_p, causal)
    out_t = model.inference(tree_tokens, pos_t, dense)[0]
    ref_noise = _rel(ref, oracle) if ref is not None else 0.0
    ours = _rel(out_t, oracle)
    print(f"\nforward parity {'awq' if awq else 'fp16'} H={shape.hidden_size}: ours {ours:.2e}, reference kernels {ref_noise:.2e} "
          f"(vs the fp32 oracle, relative to max|logit|); ours vs the exact-weight oracle {_rel(out_t, res['oracle_exact']):.2e}")
    bound = max(TOL, ref_noise)
    assert ours <= bound
    top2 = oracle.topk(2, dim=-1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * bound * oracle.abs().max()
    assert torch.equal(out_t.argmax(-1)[safe], oracle.argmax(-1)[safe])
    # the compact (prefix_len, tree_bits) form gives the same rows as the dense mask
    model.kv_offset -= N
    out_c = model.verify_forward(tree_tokens[0], pos_t[0], P,
                                 torch.from_numpy(otree.tree_bits_from_mask(tm).view(np.int32)).to(DEV))
    assert torch.equal(out_c, out_t)

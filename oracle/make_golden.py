#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE's own Python on CPU.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Run in the build container only
(``python oracle/make_golden.py``): /root/reference does not exist on the GPU box, which is why
the outputs are committed as small fixtures.

What is driven, unmodified, from /root/reference (read-only import, nothing copied):
  * umbrella.sequoia_utils.generate_sequoia_tree / successor_list_to_mask and the six shipped
    growmap JSONs                                                    -> trees.json
  * StaticSpeculationEngine.initialize/_prefill/build_tree/verify    -> accept_static.json
  * DynamicSpeculationEngine.initialize/_prefill/build_tree/verify   -> accept_dynamic.npz
  * KV_Cache / StaticKV_Cache (append, gather, pure-torch attention) -> attention.npz, inside the
    engine runs above
  * umbrella.models.model_utils.apply_rotary_pos_emb                 -> rope.npz
  * transformers' LlamaRotaryEmbedding (what llama.py:45-46 reads)   -> rope.npz (inv_freq)
  * vLLM's restatement of the AutoAWQ packer (awq_pack)              -> awq_pack.npz

Shims needed to import the reference in this image: ``awq``, ``awq_ext``,
``matplotlib`` stubs; fake LMs ("bigram tables") stand in for the HF models; CUDA-graph helpers are
replaced by their un-graphed bodies; device is "cpu".
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("awq"); _stub("awq.modules")
    _stub("awq.modules.linear", WQLinear_GEMM=type("WQLinear_GEMM", (), {}))
    _stub("awq_ext")
    for mod in ("matplotlib", "matplotlib.pyplot", "networkx"):
        try:
            __import__(mod)
        except Exception:
            _stub(mod)
    sys.path.insert(0, REF)
    import umbrella  # noqa: F401


def canon_hash(obj) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


# ---------------------------------------------------------------------------------------------
def gen_trees():
    from umbrella.sequoia_utils import generate_sequoia_tree, successor_list_to_mask
    out = {"shipped": {}, "generated": {}}
    for f in sorted(glob.glob(os.path.join(REF, "umbrella/trees/*.json"))):
        d = json.load(open(f))
        assert successor_list_to_mask(d["Successors"]) == d["mask"]
        out["shipped"][os.path.basename(f)] = {
            "Successors": d["Successors"], "size": d["size"], "depth": d["depth"],
            "roots": d["roots"], "branches": d["branches"],
            "mask_sha256": canon_hash(d["mask"]), "all_sha256": canon_hash(d)}
    for (w, dpt) in [(3, 4), (4, 6), (2, 3), (4, 8)]:
        tmp = "/tmp/_seq_%dx%d.json" % (w, dpt)
        r = generate_sequoia_tree(w, dpt, json_file=tmp)
        out["generated"]["%dx%d" % (w, dpt)] = {"Successors": r["Successors"], "all_sha256": canon_hash(r)}
    acc = [0.55, 0.22, 0.1, 0.06, 0.04, 0.03]
    r = generate_sequoia_tree(5, 5, acc=acc, json_file="/tmp/_seq_acc.json")
    out["generated"]["5x5_acc"] = {"acc": acc, "Successors": r["Successors"], "all_sha256": canon_hash(r)}
    json.dump(out, open(os.path.join(OUT, "trees.json"), "w"), separators=(",", ":"))
    print("trees.json", len(out["shipped"]), "shipped,", len(out["generated"]), "generated")


# ---------------------------------------------------------------------------------------------
class FakeCfg:
    def __init__(self, layers=2, hkv=1, hq=2, d=4, vocab=64):
        self.num_hidden_layers = layers
        self.num_key_value_heads = hkv
        self.num_attention_heads = hq
        self.hidden_size = hq * d
        self.vocab_size = vocab


class BigramLM:
    """inference() returns table[input_ids]; appends identifiable K/V rows into a REAL reference
    KV cache so that gather_kv_incremental is exercised on real cache objects."""

    def __init__(self, table, cache_cls, max_length):
        from umbrella.attn import cache as refcache
        self.config = FakeCfg(vocab=table.shape[0])
        self.table = table
        self.kv_cache = getattr(refcache, cache_cls)(self.config, max_length=max_length, device="cpu",
                                                     dtype=torch.float32)
        self.static = cache_cls == "StaticKV_Cache"
        self.device = "cpu"
        self.calls = []

    def alloc(self, **kw):
        pass

    def initialize_cuda_graph(self, *a, **k):
        pass

    def _kv_rows(self, input_ids, position_ids):
        q = input_ids.shape[1]
        c = self.config
        base = input_ids[0].float() * 1000 + position_ids[0].float()          # [q]
        k = base[:, None, None].expand(q, c.num_key_value_heads, 4).clone()
        return k, -k

    def inference(self, input_ids, position_ids, attention_mask, storage_ids):
        k, v = self._kv_rows(input_ids, position_ids)
        for layer in range(self.config.num_hidden_layers):
            kk, vv = k + layer * 0.5, v - layer * 0.5
            if self.static:
                self.kv_cache.update_kv_cache(kk.transpose(0, 1), vv.transpose(0, 1), layer, storage_ids)
            else:
                self.kv_cache.update_kv_cache(kk, vv, layer, storage_ids)
        logits = self.table[input_ids]                                         # [1, q, V]
        self.calls.append((input_ids.clone(), logits.argmax(-1)[0].clone()))
        return logits.float()

    def graph_inference(self, input_ids, storage_ids, position_ids=None, attention_mask=None):
        return self.inference(input_ids, position_ids, attention_mask, storage_ids)

    def gather_kv_incremental(self, indices, offset):
        self.kv_cache.gather_kv_incremental(indices, offset)

    def clear(self):
        self.kv_cache.clear()


def make_tables(vocab, seed, agree=0.7):
    g = torch.Generator().manual_seed(seed)
    draft = torch.randn(vocab, vocab, generator=g)
    target = draft.clone()
    flip = torch.rand(vocab, generator=g) > agree            # rows where the target disagrees
    target[flip] = torch.randn(int(flip.sum()), vocab, generator=g)
    return draft, target


def patch_engine_module(mod, draft_lm, target_lm, eos):
    class _Auto:
        @staticmethod
        def from_pretrained(model_name, offload=False, cuda_graph=False, **kw):
            return draft_lm if model_name == "draft" else target_lm
    mod.AutoModelLM = _Auto
    mod.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda name: None)
    mod.GenerationConfig = types.SimpleNamespace(
        from_pretrained=lambda name: types.SimpleNamespace(eos_token_id=eos))
    if hasattr(mod, "cuda_graph_for_sampling_argmax_gather"):
        from umbrella.speculation.speculation_utils import sampling_argmax_gather

        def no_graph(device, idx_len, num_samples, dtype, dim, index_len):
            return lambda logits, idx: sampling_argmax_gather(logits, num_samples, idx)
        mod.cuda_graph_for_sampling_argmax_gather = no_graph


def cache_digest(cache, upto):
    """Rows [0, upto) of layer 0 and the last layer, head 0, element 0 (identifies token+position)."""
    k = cache.k_cache
    if k.shape[1] == cache.num_key_value_heads and k.dim() == 4 and type(cache).__name__ == "StaticKV_Cache":
        rows = k[:, 0, :upto, 0]
    else:
        rows = k[:, :upto, 0, 0]
    return rows.tolist()


def gen_accept_static():
    import umbrella.speculation.static_speculation_engine as sse
    cases = []
    vocab, max_length = 64, 256
    for f in sorted(glob.glob(os.path.join(REF, "umbrella/trees/*.json"))):
        for seed, eos in [(0, [63]), (1, [5, 9])]:
            draft_t, target_t = make_tables(vocab, 100 + seed)
            draft = BigramLM(draft_t, "StaticKV_Cache", max_length)
            target = BigramLM(target_t, "KV_Cache", max_length)
            patch_engine_module(sse, draft, target, eos)
            eng = sse.StaticSpeculationEngine("draft", "target", dtype=torch.float32, device="cpu",
                                              growmap_path=f, max_length=max_length, safe_buffer=16)
            eng.initialize()
            g = torch.Generator().manual_seed(7 + seed)
            prompt = torch.randint(3, vocab - 2, (1, 9), generator=g)
            assert eng._prefill(prompt)
            iters = []
            for it in range(14):
                if not eng.validate_status():
                    break
                n0 = eng.num_nodes
                eng.build_tree()
                N = eng.tree_size
                spec = eng.tokens[0, n0:n0 + N].tolist()
                pos_before = eng.position_ids[0, :n0 + N].tolist()
                target.calls.clear()
                cont = eng.verify()
                sampled = target.calls[-1][1].tolist()
                n1 = eng.num_nodes
                iters.append({
                    "num_nodes_before": n0, "speculated": spec, "sampled": sampled,
                    "position_ids_before": pos_before,
                    "continue": bool(cont), "num_nodes_after": n1,
                    "tokens_after": eng.tokens[0, :n1 + 1].tolist(),
                    "position_ids_after": eng.position_ids[0, :n1 + N].tolist(),
                    "target_kv_offset": target.kv_cache.kv_offset,
                    "draft_kv_offset": draft.kv_cache.kv_offset,
                    "target_k_digest": cache_digest(target.kv_cache, n1),
                    "draft_k_digest": cache_digest(draft.kv_cache, n1),
                    "target_tail_zero": bool((target.kv_cache.k_cache[:, n1:] == 0).all()),
                })
                if not cont:
                    break
            cases.append({"tree": os.path.basename(f), "seed": seed, "eos": eos, "vocab": vocab,
                          "max_length": max_length, "prompt": prompt[0].tolist(),
                          "first_token": int(iters[0]["speculated"][0]),
                          "parents": eng.parents.tolist(), "depth": eng.depth.tolist(),
                          "node_in_path": eng.node_in_path.tolist(), "iters": iters})
    # adversarial hand-scripted cases on the 3x4 tree (EOS at root / mid-path / as bonus,
    # root-only accept, full-depth accept, duplicate tokens under different parents)
    json.dump(cases, open(os.path.join(OUT, "accept_static.json"), "w"), separators=(",", ":"))
    print("accept_static.json", len(cases), "cases,", sum(len(c["iters"]) for c in cases), "iterations")


def gen_accept_scripted():
    """Reference accept block driven directly with scripted sampled/speculated vectors (the exact
    statements of static_speculation_engine.py:312-349 executed through the real verify())."""
    import umbrella.speculation.static_speculation_engine as sse
    from umbrella.attn.cache import KV_Cache

    class ScriptLM:
        def __init__(self, vocab, max_length):
            self.config = FakeCfg(vocab=vocab)
            self.kv_cache = KV_Cache(self.config, max_length=max_length, device="cpu", dtype=torch.float32)
            self.next_sampled = None

        def alloc(self, **kw):
            pass

        def initialize_cuda_graph(self, *a, **k):
            pass

        def inference(self, input_ids, position_ids, attention_mask, storage_ids):
            q = input_ids.shape[1]
            rows = (position_ids[0].float() * 100 + torch.arange(q).float())[:, None, None].expand(q, 1, 4)
            for layer in range(2):
                self.kv_cache.update_kv_cache(rows + layer, -rows, layer, storage_ids)
            logits = torch.zeros(1, q, self.config.vocab_size)
            logits[0, torch.arange(q), self.next_sampled] = 1.0
            return logits

        def gather_kv_incremental(self, indices, offset):
            self.kv_cache.gather_kv_incremental(indices, offset)

        def clear(self):
            self.kv_cache.clear()

    f = os.path.join(REF, "umbrella/trees/sequoia_tree-3x4.json")
    succ = json.load(open(f))["Successors"]
    vocab, max_length, n0 = 200, 128, 20
    EOS = [199]
    # speculated tokens: node i carries token 100+i unless overridden
    scripts = {
        "root_only": dict(sampled={0: 50}),
        "full_depth": dict(sampled={0: 101, 1: 104, 4: 107, 7: 110, 10: 77}),
        "second_child": dict(sampled={0: 102, 2: 106, 6: 109, 9: 3}),
        "eos_at_root": dict(spec={0: 199}, sampled={0: 101, 1: 104}),
        "eos_mid_path": dict(spec={4: 199}, sampled={0: 101, 1: 199, 4: 107, 7: 110}),
        "eos_as_bonus": dict(sampled={0: 101, 1: 199}),
        "dup_tokens_other_parent": dict(spec={6: 104}, sampled={0: 102, 2: 104, 1: 104}),
        "third_child_leaf": dict(sampled={0: 103, 3: 150}),
    }
    out = []
    for name, sc in scripts.items():
        draft, target = ScriptLM(vocab, max_length), ScriptLM(vocab, max_length)
        patch_engine_module(sse, draft, target, EOS)
        eng = sse.StaticSpeculationEngine("draft", "target", dtype=torch.float32, device="cpu",
                                          growmap_path=f, max_length=max_length, safe_buffer=16)
        eng.initialize()
        N = eng.tree_size
        # emulate a committed prefix of n0 tokens whose KV is in both caches
        eng.num_nodes = n0
        eng.tokens[0, :n0] = torch.arange(1, n0 + 1)
        eng.position_ids[0, :n0] = torch.arange(n0)
        eng.position_ids[0, n0:n0 + N] = n0 + eng.depth
        for lm in (draft, target):
            lm.next_sampled = torch.zeros(n0, dtype=torch.long)
            lm.inference(eng.tokens[:, :n0], eng.position_ids[:, :n0], None, eng.storage_ids[:n0])
        eng.num_nodes_this_iter = n0 + N
        eng.attn_mask_this_iter = eng.attn_mask[max_length - n0 - N: max_length,
                                                max_length - n0 - N: 2 * max_length - n0 - N].contiguous()
        spec = [100 + i for i in range(N)]
        for k, v in sc.get("spec", {}).items():
            spec[int(k)] = v
        sampled = [1] * N
        for k, v in sc["sampled"].items():
            sampled[int(k)] = v
        eng.tokens[0, n0:n0 + N] = torch.tensor(spec)
        eng.num_draft_tokens_this_iter = n0 + N
        target.next_sampled = torch.tensor(sampled)
        draft.next_sampled = torch.zeros(N, dtype=torch.long)
        # the draft cache also holds the tree rows (build_tree would have put them there)
        draft.inference(eng.tokens[:, n0:n0 + N], eng.position_ids[:, n0:n0 + N], None, eng.storage_ids[n0:n0 + N])
        cont = eng.verify()
        n1 = eng.num_nodes
        out.append({"name": name, "num_nodes_before": n0, "speculated": spec, "sampled": sampled, "eos": EOS,
                    "continue": bool(cont), "num_nodes_after": n1,
                    "tokens_after": eng.tokens[0, :n1 + 1].tolist(),
                    "position_ids_after": eng.position_ids[0, :n1 + N].tolist(),
                    "target_k_digest": target.kv_cache.k_cache[:, :n1, 0, 0].tolist(),
                    "target_kv_offset": target.kv_cache.kv_offset})
    json.dump({"tree": "sequoia_tree-3x4.json", "Successors": succ, "cases": out},
              open(os.path.join(OUT, "accept_scripted.json"), "w"), separators=(",", ":"))
    print("accept_scripted.json", len(out), "cases")


def gen_accept_dynamic():
    import umbrella.speculation.dynamic_speculation_engine as dse
    vocab, max_length = 64, 512
    runs = {}
    for ci, (width, beams, depth, seed, eos) in enumerate([(4, 6, 3, 0, [63]), (8, 8, 5, 1, [7]), (16, 24, 6, 2, [63])]):
        draft_t, target_t = make_tables(vocab, 200 + seed, agree=0.8)
        draft = BigramLM(draft_t, "KV_Cache", max_length)
        target = BigramLM(target_t, "KV_Cache", max_length)
        patch_engine_module(dse, draft, target, eos)
        eng = dse.DynamicSpeculationEngine("draft", "target", dtype=torch.float32, device="cpu",
                                           max_length=max_length, safe_buffer=16, width=width,
                                           num_beams=beams, depth=depth)
        eng.initialize()
        g = torch.Generator().manual_seed(17 + seed)
        prompt = torch.randint(3, vocab - 2, (1, 11), generator=g)
        assert eng._prefill(prompt)
        N = eng.tree_size
        rec = {k: [] for k in ("n0", "spec", "parents", "tree_mask", "sampled", "cont", "n1", "tokens_after",
                               "pos_after", "k_digest")}
        for it in range(10):
            if not eng.validate_status():
                break
            n0 = eng.num_nodes
            eng.build_tree()
            rec["n0"].append(n0)
            rec["spec"].append(eng.tokens[0, n0:n0 + N].numpy().copy())
            rec["parents"].append(eng.parents.numpy().copy())
            rec["tree_mask"].append(eng.attn_mask_this_iter[n0:n0 + N, n0:n0 + N].numpy().copy())
            target.calls.clear()
            cont = eng.verify()
            rec["sampled"].append(target.calls[-1][1].numpy().copy())
            n1 = eng.num_nodes
            rec["cont"].append(bool(cont)); rec["n1"].append(n1)
            ta = np.zeros(max_length, dtype=np.int64); ta[:n1 + 1] = eng.tokens[0, :n1 + 1].numpy()
            pa = np.zeros(max_length, dtype=np.int64); pa[:n1 + N] = eng.position_ids[0, :n1 + N].numpy()
            kd = np.zeros(max_length, dtype=np.float32); kd[:n1] = target.kv_cache.k_cache[0, :n1, 0, 0].numpy()
            rec["tokens_after"].append(ta); rec["pos_after"].append(pa); rec["k_digest"].append(kd)
            assert draft.kv_cache.kv_offset == target.kv_cache.kv_offset == n1
            if not cont:
                break
        pre = "c%d_" % ci
        runs[pre + "meta"] = np.array([width, beams, depth, vocab, max_length, len(rec["n0"])] + eos, dtype=np.int64)
        runs[pre + "prompt"] = prompt[0].numpy()
        runs[pre + "draft_table"] = draft_t.numpy()
        runs[pre + "target_table"] = target_t.numpy()
        runs[pre + "depth"] = eng.depth.numpy()
        for k, v in rec.items():
            runs[pre + k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, "accept_dynamic.npz"), **runs)
    print("accept_dynamic.npz", len(runs), "arrays")


# ---------------------------------------------------------------------------------------------
def gen_attention():
    """StaticKV_Cache.compute_attention (cache.py:169-192) in fp32 on CPU = ground truth for the
    masked GQA attention; inputs are fp16-representable so the same vectors feed the fp16 kernel."""
    from umbrella.attn.cache import StaticKV_Cache
    out = {}
    for ci, (hq, hkv, d, prefix, tree_file, seed) in enumerate([
            (8, 2, 16, 19, "sequoia_tree-3x4.json", 0),
            (8, 1, 32, 40, "8b_sequoia_tree-5x6.json", 1),
            (4, 4, 16, 5, "sequoia_tree-3x4.json", 2)]):
        tree = json.load(open(os.path.join(REF, "umbrella/trees", tree_file)))
        tmask = torch.tensor(tree["mask"]) == 1
        n = tree["size"]
        s = prefix + n
        cfg = FakeCfg(layers=1, hkv=hkv, hq=hq, d=d)
        cache = StaticKV_Cache(cfg, max_length=s, device="cpu", dtype=torch.float32)
        g = torch.Generator().manual_seed(50 + seed)
        q = torch.randn(1, hq, n, d, generator=g).half().float()
        k_all = torch.randn(hkv, s, d, generator=g).half().float()
        v_all = torch.randn(hkv, s, d, generator=g).half().float()
        cache.k_cache[0, :, :prefix] = k_all[:, :prefix]
        cache.v_cache[0, :, :prefix] = v_all[:, :prefix]
        mask = torch.zeros(n, s, dtype=torch.bool)
        mask[:, :prefix] = True
        mask[:, prefix:] = tmask
        storage = torch.arange(prefix, s)
        o = cache.compute_attention(q, k_all[None, :, prefix:], v_all[None, :, prefix:], 0, storage, mask)
        pre = "c%d_" % ci
        out[pre + "meta"] = np.array([hq, hkv, d, prefix, n], dtype=np.int64)
        out[pre + "q"] = q[0].transpose(0, 1).contiguous().numpy()          # [n, hq, d]
        out[pre + "k"] = k_all.transpose(0, 1).contiguous().numpy()         # [s, hkv, d]
        out[pre + "v"] = v_all.transpose(0, 1).contiguous().numpy()
        out[pre + "tree_mask"] = tmask.numpy()
        out[pre + "out"] = o[0].reshape(n, hq, d).numpy()                   # [n, hq, d]
    np.savez_compressed(os.path.join(OUT, "attention.npz"), **out)
    print("attention.npz", len(out), "arrays")


def gen_rope():
    from umbrella.models.model_utils import apply_rotary_pos_emb
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding
    out = {}
    cfgs = {
        "70b": dict(hidden_size=8192, num_attention_heads=64, num_key_value_heads=8, rope_theta=500000.0,
                    max_position_embeddings=131072,
                    rope_scaling={"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                  "original_max_position_embeddings": 8192, "rope_type": "llama3"}),
        "1b": dict(hidden_size=2048, num_attention_heads=32, num_key_value_heads=8, rope_theta=500000.0,
                   max_position_embeddings=131072,
                   rope_scaling={"factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                 "original_max_position_embeddings": 8192, "rope_type": "llama3"}),
    }
    for name, kw in cfgs.items():
        try:
            cfg = LlamaConfig(num_hidden_layers=1, intermediate_size=64, vocab_size=32, **kw)
        except TypeError:
            rp = dict(kw.pop("rope_scaling")); rp["rope_theta"] = kw.pop("rope_theta")
            cfg = LlamaConfig(num_hidden_layers=1, intermediate_size=64, vocab_size=32, rope_parameters=rp, **kw)
        rot = LlamaRotaryEmbedding(config=cfg)
        out[name + "_inv_freq"] = rot.inv_freq.detach().numpy()
        out[name + "_attention_scaling"] = np.array([rot.attention_scaling], dtype=np.float64)
    # apply_rotary_pos_emb in fp16 on CPU, tables as llama.py:47-60 builds them
    inv_freq = torch.from_numpy(out["70b_inv_freq"])
    S, D = 96, 128
    pos = torch.arange(0, S).unsqueeze(0)
    freqs = (inv_freq[None, :, None].float().expand(1, -1, 1) @ pos[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[0].half(), emb.sin()[0].half()
    g = torch.Generator().manual_seed(3)
    q = torch.randn(1, 13, 4, D, generator=g).half()
    k = torch.randn(1, 13, 2, D, generator=g).half()
    position_ids = torch.tensor([[40, 41, 41, 41, 42, 42, 42, 43, 43, 43, 44, 44, 44]])
    qe, ke = apply_rotary_pos_emb(q, k, cos, sin, position_ids)
    out.update(rope_cos=cos.numpy(), rope_sin=sin.numpy(), rope_q=q.numpy(), rope_k=k.numpy(),
               rope_pos=position_ids.numpy(), rope_q_out=qe.numpy(), rope_k_out=ke.numpy())
    np.savez_compressed(os.path.join(OUT, "rope.npz"), **out)
    print("rope.npz", len(out), "arrays")


def gen_awq_pack():
    from vllm.model_executor.layers.quantization.utils.quant_utils import awq_pack
    rng = np.random.default_rng(11)
    out = {}
    for i, (k, n) in enumerate([(4, 8), (32, 64), (128, 256)]):
        q = rng.integers(0, 16, size=(k, n), dtype=np.int64)
        packed = awq_pack(torch.from_numpy(q).to(torch.int32), 4, k, n).numpy()
        out["q%d" % i] = q.astype(np.uint8)
        out["packed%d" % i] = packed
    np.savez_compressed(os.path.join(OUT, "awq_pack.npz"), **out)
    print("awq_pack.npz", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import_reference()
    torch.set_num_threads(4)
    gen_trees()
    gen_accept_static()
    gen_accept_scripted()
    gen_accept_dynamic()
    gen_attention()
    gen_rope()
    gen_awq_pack()

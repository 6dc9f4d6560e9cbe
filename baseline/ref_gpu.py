#!/usr/bin/env python
"""The reference's own GPU path timed on this box -- MEASUREMENT HARNESS ONLY.

Runs the UNMODIFIED reference package from baseline/_ref (``pip install --no-index --no-deps --target baseline/_ref
<copy of /root/reference>``, git-ignored, travels to the GPU box) through its own public classes
(``StaticSpeculationEngine`` / ``AutoModelLM`` / ``LlamaAwq`` / ``LlamaCudagraph`` / ``KV_Cache``) on the same synthetic
weights, prompt, tree and scripted acceptance as bench.py's own arm.  Nothing of umbrella_b200's engine, models or
kernels is on the timed path; from this repo only the synthetic-weight generator, the growmap data and the scripted
acceptance are used, so both arms see the same inputs.

The reference cannot run as-is in this image; the version shims below stand in for absent / drifted third-party
packages, none of them touches reference source:
  1. ``awq`` (autoawq 0.2.7.post3) absent              -> stub module providing the ``WQLinear_GEMM`` name
  2. ``awq_ext`` (autoawq-kernels 0.0.8) absent        -> vLLM 0.22's build of the same upstream CUDA kernels:
        gemm_forward_cuda(x, qweight, scales, qzeros, 8)  -> vllm._custom_ops.awq_gemm (same argument order)
        dequantize_weights_cuda(...)                       -> vllm._custom_ops.awq_dequantize
  3. ``matplotlib`` absent                                -> stub
  4. flashinfer 0.6.11: ``allow_fp16_qk_reduction`` was renamed ``use_fp16_qk_reduction`` and this build cannot compile
     the fp16-reduction kernel (prefill.cuh:182) -> the kwarg is dropped (fp32 QK reduction: slower, if anything)
  5. transformers 5.5: ``LlamaConfig.rope_theta`` moved; no checkpoints offline -> ``*.from_pretrained`` return objects
     built from the synthetic weights (tensor holders shaped like the HF modules the reference reads from)
  6. the tokenizer / GenerationConfig need hub files     -> token-id entry points only, eos list empty

Variants: C = static engine, resident ``LlamaAwq`` target + ``LlamaCudagraph`` draft (the strongest
reference configuration on a 180 GB GPU).  A = ``LlamaAwqOffload`` with ``num_cache_layers = 0`` is reported by
bench.py from our own paged model only (the reference's copy loop needs 35.6 GB of pinned host weights).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(HERE, "fi_ws"))
os.environ.setdefault("FLASHINFER_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("VLLM_USE_TRITON_AWQ", "0")


def install_shims():
    import torch
    awq = types.ModuleType("awq")
    awq_mod = types.ModuleType("awq.modules")
    awq_lin = types.ModuleType("awq.modules.linear")

    class WQLinear_GEMM:                                   # noqa: N801  (the name the reference imports)
        pass
    awq_lin.WQLinear_GEMM = WQLinear_GEMM
    awq.modules, awq_mod.linear = awq_mod, awq_lin
    sys.modules.update({"awq": awq, "awq.modules": awq_mod, "awq.modules.linear": awq_lin})

    from vllm import _custom_ops as vops
    awq_ext = types.ModuleType("awq_ext")
    awq_ext.gemm_forward_cuda = lambda x, qweight, scales, qzeros, split_k: vops.awq_gemm(x, qweight, scales, qzeros, split_k)
    awq_ext.dequantize_weights_cuda = lambda qweight, scales, qzeros, a, b, c, d: vops.awq_dequantize(qweight, scales, qzeros, a, b, c)
    sys.modules["awq_ext"] = awq_ext

    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})

    import flashinfer
    orig = flashinfer.single_prefill_with_kv_cache

    def single_prefill_with_kv_cache(*args, allow_fp16_qk_reduction=None, **kw):
        return orig(*args, **kw)
    flashinfer.single_prefill_with_kv_cache = single_prefill_with_kv_cache
    if not hasattr(flashinfer.sampling, "top_p_renorm_prob"):
        flashinfer.sampling.top_p_renorm_prob = flashinfer.sampling.top_p_renorm_probs
    return torch


class Holder:
    """A tensor slot shaped like ``module.weight``: ``.detach()`` / ``.pin_memory()`` hand the device tensor through (the
    synthetic weights are generated on the GPU; the reference's pin-then-copy staging has nothing to stage)."""

    def __init__(self, t):
        self.t = t

    def detach(self):
        return self

    def pin_memory(self):
        return self

    def to(self, device, non_blocking=False):
        return self.t.to(device)

    def copy_(self, *a, **k):
        raise RuntimeError("Holder is read-only")


def fake_hf_model(shape, awq: bool, seed: int, device: str, torch, L):
    """An object exposing exactly what the reference's alloc() reads off a Hugging Face model (llama.py:39-73,228-258,
    421-457; llama_layer.py:25-40,105-124,188-203; awq_utils.py:18-30), filled with bench.py's synthetic weights."""
    ns = types.SimpleNamespace
    head = L.synth_head(shape, seed, device)
    inv_freq, scaling = L.llama3_inv_freq(shape)

    def lin(w):
        if awq:
            qw, qz, sc = w
            return ns(qweight=Holder(qw), qzeros=Holder(qz), scales=Holder(sc), bias=None, w_bit=4, group_size=shape.group_size,
                      in_features=qw.shape[0], out_features=sc.shape[1])
        return ns(weight=w)

    class Layers(list):
        pass
    layers = Layers()
    for i in range(shape.num_hidden_layers):
        w = L.synth_layer(shape, i, awq, seed, device)
        norm = (lambda t: Holder(t)) if awq else (lambda t: t)
        layers.append(ns(self_attn=ns(q_proj=lin(w["wq"]), k_proj=lin(w["wk"]), v_proj=lin(w["wv"]), o_proj=lin(w["wo"])),
                         mlp=ns(gate_proj=lin(w["gate_proj"]), up_proj=lin(w["up_proj"]), down_proj=lin(w["down_proj"])),
                         input_layernorm=ns(weight=norm(w["input_layernorm"]), variance_epsilon=shape.rms_norm_eps),
                         post_attention_layernorm=ns(weight=norm(w["post_attention_layernorm"]),
                                                     variance_epsilon=shape.rms_norm_eps)))
    model = ns(embed_tokens=ns(weight=head["embed_tokens"]), norm=ns(weight=head["norm"], variance_epsilon=shape.rms_norm_eps),
               rotary_emb=ns(inv_freq=inv_freq.to(device), attention_scaling=scaling), layers=layers)
    lm_head = head["embed_tokens"] if head["lm_head"] is None else head["lm_head"]
    return ns(model=model, lm_head=ns(weight=lm_head))


def hf_config(shape, transformers):
    cfg = transformers.LlamaConfig(hidden_size=shape.hidden_size, intermediate_size=shape.intermediate_size,
                                   num_hidden_layers=shape.num_hidden_layers, num_attention_heads=shape.num_attention_heads,
                                   num_key_value_heads=shape.num_key_value_heads, vocab_size=shape.vocab_size,
                                   rms_norm_eps=shape.rms_norm_eps, max_position_embeddings=shape.max_position_embeddings,
                                   tie_word_embeddings=shape.tie_word_embeddings)
    cfg.rope_theta = shape.rope_theta            # shim 5: attribute the reference reads (llama.py:32)
    cfg.eos_token_id = list(shape.eos_token_id)
    return cfg


def build_reference_engine(device: str, tree: str, target_name: str, draft_name: str, max_length: int):
    torch = install_shims()
    sys.path.insert(0, REF)                      # the reference package shadows this repo's `umbrella` alias
    sys.path.insert(1, ROOT)
    import transformers
    from umbrella_b200.models import llama as L
    from umbrella_b200.sequoia_utils import load_growmap
    import umbrella                              # noqa: F401  -> baseline/_ref/umbrella
    assert os.path.realpath(os.path.dirname(umbrella.__file__)).startswith(os.path.realpath(REF)), umbrella.__file__
    from umbrella.models import llama as ref_llama
    from umbrella.speculation import static_speculation_engine as ref_static

    shapes = {name: L.KNOWN_MODELS[name] for name in (target_name, draft_name)}
    seeds = {target_name: 1234, draft_name: 4321}
    transformers.LlamaConfig.from_pretrained = classmethod(lambda cls, name, **kw: hf_config(shapes[name][0], transformers))

    def from_pretrained(name, **kw):
        shape, awq = shapes[name]
        return fake_hf_model(shape, awq, seeds[name], device, torch, L)
    for mod in (ref_llama,):
        mod.AutoModelForCausalLM = types.SimpleNamespace(from_pretrained=from_pretrained)
        mod.LlamaForCausalLM = types.SimpleNamespace(from_pretrained=from_pretrained)
    ref_static.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda *a, **k: types.SimpleNamespace(
        decode=lambda *a, **k: "", encode=lambda *a, **k: None))
    ref_static.GenerationConfig = types.SimpleNamespace(from_pretrained=lambda *a, **k: types.SimpleNamespace(eos_token_id=[]))

    gm = load_growmap(tree)
    gm_path = os.path.join(HERE, "_ref", f"growmap_{os.path.basename(tree)}")
    json.dump({k: gm[k] for k in ("roots", "branches", "Successors", "mask", "depth", "size")}, open(gm_path, "w"))
    eng = ref_static.StaticSpeculationEngine(draft_model_name=draft_name, target_model_name=target_name, device=device,
                                             dtype=torch.float16, growmap_path=gm_path, max_length=max_length, temperature=0.0)
    eng.initialize()
    eng.eos_tokens = []
    return eng, torch, gm


def attach_script(eng, torch, seed: int):
    """Scripted acceptance without touching the reference's verify(): after the real target forward over the tree, the
    logit of the token the script declares the target's choice is lifted above the row maximum."""
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    script = ScriptedAcceptance(eng.Successors, seed=seed, max_iters=4096, device=eng.device)
    inner = eng.target_model.inference

    def inference(*a, **kw):
        logits = inner(*a, **kw)
        if logits.shape[1] == eng.tree_size and eng.num_draft_tokens_this_iter > eng.num_nodes:
            sampled = logits[0].argmax(dim=-1)
            chosen = script(eng, sampled)
            logits[0].scatter_(1, chosen[:, None], float(logits.max()) + 1.0)
        return logits
    eng.target_model.inference = inference
    return script


def run(args):
    dev = "cuda:0"
    t_setup = time.time()
    eng, torch, gm = build_reference_engine(dev, args.tree, args.target, args.draft, args.max_length)
    attach_script(eng, torch, seed=100)
    setup_s = time.time() - t_setup
    g = torch.Generator().manual_seed(0)
    prompt = torch.randint(3, 30000, (1, args.prompt_len), generator=g).to(dev)

    def new_request():
        eng.reset()
        assert eng._prefill(prompt)
        return eng.num_nodes
    start = new_request()
    for _ in range(max(args.warmup, 3)):
        eng.build_tree()
        eng.verify()
    torch.cuda.synchronize()
    tokens, dev_ms, steps_left = 0, 0.0, args.steps
    verify_ms, build_ms = [], []
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    while steps_left > 0:
        if eng.num_nodes - start >= args.gen_len or not eng.validate_status():
            start = new_request()
        n0 = eng.num_nodes
        e0.record()
        eng.build_tree()
        e1.record()
        eng.verify()
        e2.record()
        torch.cuda.synchronize()
        build_ms.append(e0.elapsed_time(e1))
        verify_ms.append(e1.elapsed_time(e2))
        dev_ms += e0.elapsed_time(e2)
        tokens += eng.num_nodes - n0
        steps_left -= 1
    # end to end through the reference's public generate()
    t0 = time.time()
    out = eng.generate(input_ids=prompt[0].tolist(), max_new_tokens=args.gen_len)
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    print(json.dumps({
        "impl": "reference_gpu", "variant": "C: StaticSpeculationEngine, resident LlamaAwq target + LlamaCudagraph draft",
        "value": tokens / (dev_ms * 1e-3), "unit": "tokens/s", "steps": args.steps, "ms_per_step": dev_ms / args.steps,
        "ms_per_verify": sum(verify_ms) / len(verify_ms), "ms_per_build_tree": sum(build_ms) / len(build_ms),
        "mean_accept_len": tokens / args.steps, "e2e_tokens_per_s": len(out["generated_tokens"]) / e2e_s,
        "setup_s": setup_s, "cpu_count": os.cpu_count(),
        "kernels": "vLLM 0.22 _C.awq_gemm (upstream AWQ kernel, split_k 8), flashinfer 0.6.11 single_prefill_with_kv_cache("
                   "custom_mask, fp32 QK reduction) + rmsnorm, torch F.linear / CUDA-graphed draft",
        "workload": f"{args.target} + {args.draft}, {args.tree}, prompt {args.prompt_len}, generation {args.gen_len}, "
                    f"max_length {args.max_length}, same synthetic weights / prompt / scripted acceptance as the product arm"}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tree", default="sequoia_tree-3x4.json")
    ap.add_argument("--target", default="hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4")
    ap.add_argument("--draft", default="meta-llama/Llama-3.2-1B-Instruct")
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--gen-len", type=int, default=256)
    ap.add_argument("--max-length", type=int, default=2048)
    run(ap.parse_args())

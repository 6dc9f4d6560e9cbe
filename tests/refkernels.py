"""The reference's OWN GPU kernels, op by op -- the second comparator of the forward parity tests.

TEST INFRASTRUCTURE ONLY (GPU box).  The reference forward (umbrella/models/llama.py:262-322) is plain
torch glue around three third-party kernels; this module restates that glue and calls the same kernels:

  * ``AwqLinear.apply`` (umbrella/quantization/awq_utils.py:63-86) -> ``awq_ext.gemm_forward_cuda(x, qweight,
    scales, qzeros, 8)``.  ``awq_ext`` (autoawq-kernels 0.0.8) is absent from the image; vLLM 0.22 ships a build
    of the same upstream CUDA source behind ``vllm._custom_ops.awq_gemm`` (same argument order, split_k_iters = 8).
  * ``layer_norm`` (umbrella/models/model_utils.py:54-64) -> ``flashinfer.rmsnorm``.
  * ``KV_Cache.compute_attention`` (umbrella/attn/cache.py:67-87) -> ``flashinfer.single_prefill_with_kv_cache(
    q, k, v, kv_layout="NHD", custom_mask=mask, allow_fp16_qk_reduction=True)``.  The kwarg is called
    ``use_fp16_qk_reduction`` in flashinfer 0.6.11, and this image's flashinfer cannot build the fp16-reduction
    variant at all (static_assert "install boost_math", prefill.cuh:182), so the fp32-reduction kernel -- the more
    accurate one -- stands in; this is stated wherever the numbers are quoted.
  * RoPE, SiLU, residual adds, lm_head: the reference's own torch fp16 ops, restated verbatim.

flashinfer JIT-compiles its modules; they are pre-built for sm_100a into baseline/fi_ws (git-ignored, travels to
the GPU box) by scripts/build_flashinfer_modules.py, and FLASHINFER_WORKSPACE_BASE points there.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(ROOT, "baseline", "fi_ws"))
os.environ.setdefault("FLASHINFER_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("VLLM_USE_TRITON_AWQ", "0")

import torch


def available() -> str | None:
    """None when the kernels can be called, else the reason."""
    try:
        import flashinfer  # noqa: F401
        from vllm import _custom_ops  # noqa: F401
        if not hasattr(torch.ops._C, "awq_gemm"):
            return "vLLM _C.awq_gemm missing"
    except Exception as e:  # pragma: no cover
        return repr(e)
    return None


def awq_linear(x: torch.Tensor, w) -> torch.Tensor:
    """AwqLinear.apply for fewer than 1024 rows (awq_utils.py:74-77); fp16 models: F.linear (llama.py:87-110)."""
    if isinstance(w, torch.Tensor):
        return torch.nn.functional.linear(x, w)
    from vllm import _custom_ops as vops
    qweight, qzeros, scales = w
    return vops.awq_gemm(x.reshape(-1, x.shape[-1]).contiguous(), qweight, scales, qzeros, 8)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    import flashinfer
    return flashinfer.rmsnorm(x.contiguous(), w, eps)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids):
    """model_utils.py:24-52 on [n, H, D] tensors (unsqueeze_dim = the head axis)."""
    c = cos[position_ids].unsqueeze(1)
    s = sin[position_ids].unsqueeze(1)
    return (q * c) + (rotate_half(q) * s), (k * c) + (rotate_half(k) * s)


def attention(q, k, v, mask) -> torch.Tensor:
    """cache.py:77-85.  q [n, Hq, D], k / v [kv, Hkv, D] fp16, mask bool [n, kv]."""
    import flashinfer
    return flashinfer.single_prefill_with_kv_cache(q.contiguous(), k.contiguous(), v.contiguous(), kv_layout="NHD",
                                                   custom_mask=mask.contiguous(), use_fp16_qk_reduction=False)


def layer_forward(shape, raw: dict, hidden, position_ids, mask, k_layer, v_layer, kv_before: int, cos, sin):
    """LlamaAwq.layer_compute (llama.py:262-303) over the raw AWQ triplets of one layer.  hidden: [n, H] fp16."""
    n = hidden.shape[0]
    hq, hkv, d = shape.num_attention_heads, shape.num_key_value_heads, shape.head_dim
    residual = hidden
    x = rmsnorm(hidden, raw["input_layernorm"], shape.rms_norm_eps)
    q = awq_linear(x, raw["wq"]).view(n, hq, d)
    k = awq_linear(x, raw["wk"]).view(n, hkv, d)
    v = awq_linear(x, raw["wv"]).view(n, hkv, d)
    q, k = apply_rotary_pos_emb(q, k, cos, sin, position_ids)
    k_layer[kv_before:kv_before + n] = k
    v_layer[kv_before:kv_before + n] = v
    kv = kv_before + n
    a = attention(q, k_layer[:kv], v_layer[:kv], mask[:, :kv]).reshape(n, hq * d)
    hidden = residual + awq_linear(a, raw["wo"])
    residual = hidden
    x = rmsnorm(hidden, raw["post_attention_layernorm"], shape.rms_norm_eps)
    up = awq_linear(x, raw["up_proj"])
    gate = torch.nn.functional.silu(awq_linear(x, raw["gate_proj"]))
    hidden = residual + awq_linear(gate * up, raw["down_proj"])
    return hidden


def final_logits(shape, hidden, norm_weight, lm_head):
    """llama.py:316-322."""
    return torch.nn.functional.linear(rmsnorm(hidden, norm_weight, shape.rms_norm_eps), lm_head).float()


def three_way_logits(L, ollama, shape, oshape, awq: bool, seed: int, layers: int, prompt, tree_tokens, pos_p, pos_t, causal,
                     dense, cos, sin, device, on_layer=None, max_length: int = 256):
    """Stream ``layers`` synthetic decoder layers (models/llama.py synth_layer, generated on ``device``) through three
    implementations -- the fp32 oracle with the reference's fp16 rounding points and fp16-rounded weights, the same with
    exact (unrounded) weights, and the reference's own kernels -- over a causal prefill of the prompt rows followed by the
    tree rows.  ``on_layer(l, raw)`` lets the caller feed the same tensors to the product model.  Returns the tree-row
    logits {"oracle", "oracle_exact", "reference" (None when the kernels are unavailable)} and the head tensors."""
    head = L.synth_head(shape, seed, gen_device=device)
    lm_head = head["embed_tokens"] if head["lm_head"] is None else head["lm_head"]
    why = available()
    emb_p = torch.nn.functional.embedding(prompt, head["embed_tokens"])
    emb_t = torch.nn.functional.embedding(tree_tokens, head["embed_tokens"])
    h16, hx, hr = [emb_p.clone(), emb_t.clone()], [emb_p.clone(), emb_t.clone()], [emb_p[0].clone(), emb_t[0].clone()]
    hkv, d, P = shape.num_key_value_heads, shape.head_dim, prompt.shape[1]
    for l in range(layers):
        raw = L.synth_layer(shape, l, awq, seed, gen_device=device)
        if on_layer is not None:
            on_layer(l, raw)
        for hs, exact in ((h16, False), (hx, True)):
            w = ollama.oracle_layer_weights(raw, 128, awq, exact=exact)
            k_l = torch.zeros(max_length, hkv, d, dtype=torch.float16, device=device)
            v_l = torch.zeros_like(k_l)
            hs[0] = ollama.layer_forward(oshape, w, hs[0], pos_p, causal, k_l, v_l, 0, cos, sin)
            hs[1] = ollama.layer_forward(oshape, w, hs[1], pos_t, dense, k_l, v_l, P, cos, sin)
            del w
        if why is None:
            k_l = torch.zeros(max_length, hkv, d, dtype=torch.float16, device=device)
            v_l = torch.zeros_like(k_l)
            hr[0] = layer_forward(oshape, raw, hr[0], pos_p[0], causal, k_l, v_l, 0, cos, sin)
            hr[1] = layer_forward(oshape, raw, hr[1], pos_t[0], dense, k_l, v_l, P, cos, sin)
        del raw
    out = {"oracle": ollama.final_logits(oshape, h16[1], head["norm"], lm_head)[0],
           "oracle_exact": ollama.final_logits(oshape, hx[1], head["norm"], lm_head)[0],
           "reference": final_logits(oshape, hr[1], head["norm"], lm_head) if why is None else None,
           "reference_unavailable": why, "head": head}
    return out

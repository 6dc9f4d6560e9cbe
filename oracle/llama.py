"""Oracle: the Llama verify forward on CPU (torch, fp32 maths with the reference's fp16 rounding
points), plus a seeded synthetic-weight generator shared by tests and bench.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (paths relative to /root/reference):
  * ``layer_norm`` -> flashinfer.rmsnorm          umbrella/models/model_utils.py:54-64
  * ``apply_rotary_pos_emb`` / ``rotate_half``    umbrella/models/model_utils.py:17-52
  * cos/sin tables                               umbrella/models/llama.py:237-249
  * ``LlamaAwq.layer_compute`` / ``inference``   umbrella/models/llama.py:262-322
  * ``KV_Cache`` append + masked attention       umbrella/attn/cache.py:53-87
  * ``gather_kv_incremental``                    umbrella/attn/cache.py:41-49

Rounding points kept: the residual stream, q/k/v, RoPE products, attention output, every linear
output, SiLU and the gate*up product are fp16 tensors exactly where the reference has fp16
tensors; inside an op the maths is fp32 (fp64 for the softmax is unnecessary here).  Linear
layers use ``fp16((q - z) * s)`` weights (what both upstream AWQ kernels produce) and one rounding
of the fp32 dot product.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import awq as awq_oracle


@dataclass
class LlamaShape:
    """The subset of LlamaConfig the path reads (llama.py:24-33)."""
    hidden_size: int = 8192
    intermediate_size: int = 28672
    num_hidden_layers: int = 80
    num_attention_heads: int = 64
    num_key_value_heads: int = 8
    vocab_size: int = 128256
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_position_embeddings: int = 131072
    # llama3 rope scaling (Llama-3.1/3.2/3.3); None = plain RoPE
    rope_scaling: dict | None = field(default_factory=lambda: {
        "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
        "original_max_position_embeddings": 8192, "rope_type": "llama3"})
    tie_word_embeddings: bool = False
    group_size: int = 128

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads


LLAMA31_70B = LlamaShape()
LLAMA31_8B = LlamaShape(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                        num_attention_heads=32, num_key_value_heads=8)
LLAMA32_1B = LlamaShape(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
                        num_attention_heads=32, num_key_value_heads=8, tie_word_embeddings=True,
                        rope_scaling={"factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                      "original_max_position_embeddings": 8192, "rope_type": "llama3"})


def llama3_inv_freq(shape: LlamaShape) -> tuple[torch.Tensor, float]:
    """inv_freq and attention_scaling as ``hf_model.model.rotary_emb`` holds them (llama.py:45-46);
    restates transformers' ``_compute_llama3_parameters`` (attention_scaling is 1.0 for llama3)."""
    d = shape.head_dim
    inv_freq = 1.0 / (shape.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    rs = shape.rope_scaling
    if rs is None:
        return inv_freq, 1.0
    factor, lo, hi, old = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"], \
        rs["original_max_position_embeddings"]
    low_wl, high_wl = old / lo, old / hi
    wavelen = 2 * math.pi / inv_freq
    inv_llama = torch.where(wavelen > low_wl, inv_freq / factor, inv_freq)
    smooth = (old / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * inv_llama / factor + smooth * inv_llama
    is_medium = ~(wavelen < high_wl) * ~(wavelen > low_wl)
    return torch.where(is_medium, smoothed, inv_llama), 1.0


def rope_tables(inv_freq: torch.Tensor, attention_scaling: float, max_length: int,
                dtype=torch.float16) -> tuple[torch.Tensor, torch.Tensor]:
    """llama.py:47-60: fp32 outer product, cos/sin, * attention_scaling, cast to fp16; [S, D]."""
    position_ids = torch.arange(0, max_length).unsqueeze(0)
    inv = inv_freq[None, :, None].float().expand(1, -1, 1)
    pos = position_ids[:, None, :].float()
    freqs = (inv.float() @ pos.float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos = (emb.cos()[0] * attention_scaling).to(dtype)
    sin = (emb.sin()[0] * attention_scaling).to(dtype)
    return cos, sin


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """model_utils.py:17-21."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
               position_ids: torch.Tensor, unsqueeze_dim: int = 2):
    """model_utils.py:24-52, fp16 arithmetic (three roundings per element: two products, one sum)."""
    c = cos[position_ids].unsqueeze(unsqueeze_dim)
    s = sin[position_ids].unsqueeze(unsqueeze_dim)
    return (q * c) + (rotate_half(q) * s), (k * c) + (rotate_half(k) * s)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """flashinfer.rmsnorm semantics (model_utils.py:62): fp32 mean of squares over the row,
    ``x * rsqrt(ms + eps) * w`` in fp32, one rounding to the input dtype."""
    x32 = x.float()
    r = torch.rsqrt(x32.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (x32 * r * w.float()).to(x.dtype)


# --------------------------------------------------------------------------------------
# synthetic weights: seeded, real shapes, value-independent timing
# --------------------------------------------------------------------------------------
LINEAR_NAMES = ("wq", "wk", "wv", "wo", "gate_proj", "up_proj", "down_proj")


def linear_dims(shape: LlamaShape) -> dict[str, tuple[int, int]]:
    """(in_features K, out_features N) per linear of one decoder layer (llama.py:276-300)."""
    h, i = shape.hidden_size, shape.intermediate_size
    kvd = shape.num_key_value_heads * shape.head_dim
    return {"wq": (h, h), "wk": (h, kvd), "wv": (h, kvd), "wo": (h, h),
            "gate_proj": (h, i), "up_proj": (h, i), "down_proj": (i, h)}


def synth_awq_linear(k: int, n: int, group: int, gen: torch.Generator, scale: float = 0.003):
    """qweight int32 [K, N/8], qzeros int32 [K/g, N/8] (uniform nibbles), scales fp16 [K/g, N]
    = scale * (0.5 + U[0,1)).  Uniform int32 words = uniform nibbles."""
    qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), generator=gen, dtype=torch.int64).to(torch.int32)
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (k // group, n // 8), generator=gen, dtype=torch.int64).to(torch.int32)
    scales = (scale * (0.5 + torch.rand((k // group, n), generator=gen))).to(torch.float16)
    return qweight, qzeros, scales


def synth_layer(shape: LlamaShape, layer_idx: int, awq: bool = True, seed: int = 1234) -> dict:
    gen = torch.Generator().manual_seed(seed + layer_idx)
    out = {}
    for name, (k, n) in linear_dims(shape).items():
        if awq:
            out[name] = synth_awq_linear(k, n, shape.group_size, gen)
        else:
            out[name] = (0.02 * torch.randn((n, k), generator=gen)).to(torch.float16)   # [out, in]
    out["input_layernorm"] = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen)).to(torch.float16)
    out["post_attention_layernorm"] = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen)).to(torch.float16)
    return out


def synth_head(shape: LlamaShape, seed: int = 1234) -> dict:
    gen = torch.Generator().manual_seed(seed - 1)
    embed = (0.02 * torch.randn((shape.vocab_size, shape.hidden_size), generator=gen)).to(torch.float16)
    lm_head = embed if shape.tie_word_embeddings else \
        (0.02 * torch.randn((shape.vocab_size, shape.hidden_size), generator=gen)).to(torch.float16)
    norm = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen)).to(torch.float16)
    return {"embed_tokens": embed, "lm_head": lm_head, "norm": norm}


def dequant_fp16(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, group: int) -> torch.Tensor:
    """[K, N] fp16 = fp16((q - z) * s), via the numpy oracle (awq.dequantize)."""
    w = awq_oracle.dequantize(qweight.numpy(), qzeros.numpy(), scales.numpy(), group, np.float16)
    return torch.from_numpy(w)



# --------------------------------------------------------------------------------------
# one decoder layer, functional and device-agnostic (the 70B-shape parity tests stream the layers
# through it on the GPU box; OracleLlama below calls it with its own caches)
# --------------------------------------------------------------------------------------
def _linear(x16: torch.Tensor, w32: torch.Tensor) -> torch.Tensor:
    return (x16.float() @ w32).to(torch.float16)


def masked_attention(shape: LlamaShape, q, k_layer, v_layer, kv_len: int, mask):
    """cache.py:67-87 with the pure-torch arithmetic of cache.py:169-192: masked GQA attention over rows
    [0, kv_len) of one layer's cache, fp32 softmax, fp16 out.  q: [1, n, Hq, D] fp16; mask: bool [n, >= kv_len]."""
    n_new = q.shape[1]
    kk = k_layer[:kv_len].float()                          # [kv, Hkv, D]
    vv = v_layer[:kv_len].float()
    hq, hkv = shape.num_attention_heads, shape.num_key_value_heads
    g = hq // hkv
    qq = q[0].float().reshape(n_new, hkv, g, -1)           # [q, Hkv, g, D]
    s = torch.einsum("qhgd,khd->hgqk", qq, kk) / math.sqrt(shape.head_dim)
    s = s.masked_fill(~mask[:, :kv_len][None, None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("hgqk,khd->qhgd", p, vv)
    return o.reshape(1, n_new, hq * shape.head_dim).to(torch.float16)


def layer_forward(shape: LlamaShape, w: dict, hidden, position_ids, attention_mask, k_layer, v_layer, kv_before: int,
                  cos_cache, sin_cache):
    """llama.py:262-303.  ``w``: fp32 [K, N] matrices per linear + the two fp16 norm vectors; ``k_layer`` /
    ``v_layer``: this layer's cache [S, Hkv, D] fp16, the new rows are appended at ``kv_before``."""
    sh = shape
    bsz, q_len, _ = hidden.shape
    residual = hidden
    x = rmsnorm(hidden, w["input_layernorm"], sh.rms_norm_eps)
    q = _linear(x, w["wq"]).view(bsz, q_len, sh.num_attention_heads, sh.head_dim)
    k = _linear(x, w["wk"]).view(bsz, q_len, sh.num_key_value_heads, sh.head_dim)
    v = _linear(x, w["wv"]).view(bsz, q_len, sh.num_key_value_heads, sh.head_dim)
    q, k = apply_rope(q, k, cos_cache, sin_cache, position_ids)
    k_layer[kv_before:kv_before + q_len] = k[0]
    v_layer[kv_before:kv_before + q_len] = v[0]
    a = masked_attention(sh, q, k_layer, v_layer, kv_before + q_len, attention_mask)
    hidden = residual + _linear(a, w["wo"])
    residual = hidden
    x = rmsnorm(hidden, w["post_attention_layernorm"], sh.rms_norm_eps)
    up = _linear(x, w["up_proj"])
    gate = torch.nn.functional.silu(_linear(x, w["gate_proj"]).float()).to(torch.float16)
    hidden = residual + _linear(gate * up, w["down_proj"])
    return hidden


def final_logits(shape: LlamaShape, hidden, norm_weight, lm_head):
    """llama.py:316-322: final RMSNorm, fp16 lm_head GEMM output cast to float."""
    hidden = rmsnorm(hidden, norm_weight, shape.rms_norm_eps)
    return (hidden.float() @ lm_head.float().t()).to(torch.float16).float()


def oracle_layer_weights(raw: dict, group: int, awq: bool, exact: bool = False) -> dict:
    """raw synth_layer() tensors (any device) -> fp32 [K, N] matrices + norm vectors for layer_forward."""
    d = {}
    for name in LINEAR_NAMES:
        d[name] = awq_oracle.dequantize_torch(*raw[name], group, exact=exact) if awq else raw[name].t().float().contiguous()
    d["input_layernorm"] = raw["input_layernorm"]
    d["post_attention_layernorm"] = raw["post_attention_layernorm"]
    return d


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class OracleLlama:
    """CPU twin of ``LlamaAwq`` / ``Llama`` (llama.py:11-142, 222-322) over synthetic weights.

    ``inference`` has the reference signature and semantics (append at kv_offset, ignore the values
    of storage_ids, return fp32 logits [1, q, V])."""

    def __init__(self, shape: LlamaShape, max_length: int, awq: bool = True, seed: int = 1234,
                 num_layers: int | None = None):
        self.shape = shape
        self.max_length = max_length
        self.num_layers = shape.num_hidden_layers if num_layers is None else num_layers
        self.awq = awq
        self.seed = seed
        head = synth_head(shape, seed)
        self.embed_tokens, self.lm_head, self.norm_weight = head["embed_tokens"], head["lm_head"], head["norm"]
        inv_freq, scaling = llama3_inv_freq(shape)
        self.cos_cache, self.sin_cache = rope_tables(inv_freq, scaling, max_length)
        self.raw_layers = [synth_layer(shape, i, awq, seed) for i in range(self.num_layers)]
        # fp32 [K, N] matrices holding the fp16-rounded weights (one-time dequant)
        self.layers = []
        for lw in self.raw_layers:
            d = {}
            for name in LINEAR_NAMES:
                if awq:
                    d[name] = dequant_fp16(*lw[name], shape.group_size).float()
                else:
                    d[name] = lw[name].t().float().contiguous()
            d["input_layernorm"] = lw["input_layernorm"]
            d["post_attention_layernorm"] = lw["post_attention_layernorm"]
            self.layers.append(d)
        hkv, dh = shape.num_key_value_heads, shape.head_dim
        self.k_cache = torch.zeros(self.num_layers, max_length, hkv, dh, dtype=torch.float16)
        self.v_cache = torch.zeros_like(self.k_cache)
        self.kv_offset = 0

    # cache.py:88-91
    def clear(self):
        self.k_cache.zero_()
        self.v_cache.zero_()
        self.kv_offset = 0

    # cache.py:41-49
    def gather_kv_incremental(self, indices, offset: int):
        idx = torch.as_tensor(indices, dtype=torch.long)
        m = len(idx)
        self.k_cache[:, offset:offset + m] = self.k_cache[:, idx]
        self.v_cache[:, offset:offset + m] = self.v_cache[:, idx]
        self.k_cache[:, offset + m:] = 0.0
        self.v_cache[:, offset + m:] = 0.0
        self.kv_offset = offset + m

    def layer_compute(self, w, layer_idx, hidden, position_ids, attention_mask):
        """llama.py:262-303 (append at kv_offset: cache.py:53-65)."""
        n_new = hidden.shape[1]
        if layer_idx == 0:
            self.kv_offset += n_new
        return layer_forward(self.shape, w, hidden, position_ids, attention_mask, self.k_cache[layer_idx],
                             self.v_cache[layer_idx], self.kv_offset - n_new, self.cos_cache, self.sin_cache)

    @torch.inference_mode()
    def inference(self, input_ids, position_ids, attention_mask, storage_ids=None):
        """llama.py:305-322 -> fp32 logits [1, q, V] (fp16 GEMM output cast to float)."""
        hidden = torch.nn.functional.embedding(input_ids, self.embed_tokens)
        for idx in range(self.num_layers):
            hidden = self.layer_compute(self.layers[idx], idx, hidden, position_ids, attention_mask)
        return final_logits(self.shape, hidden, self.norm_weight, self.lm_head)

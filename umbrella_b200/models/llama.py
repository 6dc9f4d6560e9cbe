"""Llama model drivers over the native forward (umb_llama_forward).

Host-side mirror of the reference's ``Llama`` / ``LlamaAwq`` (umbrella/models/llama.py:11-142,
222-322): same ``alloc / inference / gather_kv_incremental / clear`` surface, same tensors
(``embed_tokens, lm_head, norm_weight, cos_cache, sin_cache, kv_cache``), but the per-layer Python
(``layer_compute``) is gone: one C call enqueues the whole forward on the current stream.

Weights come from one of two sources:
  * synthetic, seeded, real shapes (there are no checkpoints and no network in this environment;
    SURVEY.md 8d) -- ``SyntheticLlama``;
  * tensors handed in by a loader (``load_awq_layer`` / ``load_fp16_layer``), the seam where
    ``LlamaAwqLayer.init_parameters`` (umbrella/models/llama_layer.py:165-218) plugs in.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import torch

from .. import _lib, ops
from ..attn.cache import KV_Cache

MAX_ROWS = 128   # rows per forward chunk (UMMA N limit of the swap-AB GEMM)


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
    rope_scaling: dict | None = field(default_factory=lambda: {
        "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
        "original_max_position_embeddings": 8192, "rope_type": "llama3"})
    tie_word_embeddings: bool = False
    group_size: int = 128
    # a list, as LlamaConfig holds it (examples/generate.py:44-46 tests ``isinstance(eos_tokens, list)``)
    eos_token_id: list = field(default_factory=lambda: [128001, 128008, 128009])

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


_LLAMA3_1B_ROPE = {"factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                   "original_max_position_embeddings": 8192, "rope_type": "llama3"}
LLAMA31_70B = LlamaShape()
LLAMA31_8B = LlamaShape(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                        num_key_value_heads=8)
LLAMA32_1B = LlamaShape(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32,
                        num_key_value_heads=8, tie_word_embeddings=True, rope_scaling=_LLAMA3_1B_ROPE)

# model names of the reference's configs/*.json -> (shape, awq)
KNOWN_MODELS = {
    "hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4": (LLAMA31_70B, True),
    "casperhansen/llama-3.3-70b-instruct-awq": (LLAMA31_70B, True),
    "KirillR/QwQ-32B-Preview-AWQ": None,
    "hugging-quants/Meta-Llama-3.1-8B-Instruct-AWQ-INT4": (LLAMA31_8B, True),
    "meta-llama/Llama-3.1-8B-Instruct": (LLAMA31_8B, False),
    "meta-llama/Meta-Llama-3.1-8B-Instruct": (LLAMA31_8B, False),
    "meta-llama/Llama-3.2-1B-Instruct": (LLAMA32_1B, False),
}


def llama3_inv_freq(shape: LlamaShape) -> tuple[torch.Tensor, float]:
    """inv_freq / attention_scaling as hf_model.model.rotary_emb holds them (llama.py:45-46)."""
    d = shape.head_dim
    inv_freq = 1.0 / (shape.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    rs = shape.rope_scaling
    if rs is None:
        return inv_freq, 1.0
    factor, lo, hi, old = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"], \
        rs["original_max_position_embeddings"]
    wavelen = 2 * math.pi / inv_freq
    inv_llama = torch.where(wavelen > old / lo, inv_freq / factor, inv_freq)
    smooth = (old / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * inv_llama / factor + smooth * inv_llama
    is_medium = ~(wavelen < old / hi) * ~(wavelen > old / lo)
    return torch.where(is_medium, smoothed, inv_llama), 1.0


def rope_tables(inv_freq: torch.Tensor, attention_scaling: float, max_length: int):
    """llama.py:47-60: fp32 outer product, cos/sin, * attention_scaling, cast to fp16 -> [S, D]."""
    pos = torch.arange(0, max_length)[None, None, :].float()
    freqs = (inv_freq[None, :, None].float() @ pos).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return (emb.cos()[0] * attention_scaling).to(torch.float16), (emb.sin()[0] * attention_scaling).to(torch.float16)


class Llama:
    """fp16-weight Llama (reference ``Llama``, llama.py:11-142) on the native forward."""
    awq = False

    def __init__(self, shape: LlamaShape, max_length: int = 256, device: str = "cuda:0", dtype=torch.float16,
                 batch_size: int = 1, model_name: str | None = None, num_layers: int | None = None):
        assert batch_size == 1 and dtype == torch.float16
        self.config = shape
        self.model_name = model_name
        self.device = torch.device(device)
        self.dtype = dtype
        self.batch_size = batch_size
        self.max_length = max_length
        self.hidden_size = shape.hidden_size
        self.num_heads = shape.num_attention_heads
        self.head_dim = shape.head_dim
        self.num_key_value_heads = shape.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.num_layers = shape.num_hidden_layers if num_layers is None else num_layers
        eos = getattr(shape, "eos_token_id", (128001, 128008, 128009))
        self.eos_tokens = list(eos) if isinstance(eos, (list, tuple)) else [int(eos)]
        self.norm_variance_epsilon = shape.rms_norm_eps
        self._c_model = None
        self.layer_weights: list[dict] = []

    # ------------------------------------------------------------------ allocation
    def alloc(self, **kwargs):
        """Allocate the KV cache, RoPE tables, activations and workspaces (llama.py:35-73).  Weights
        are attached afterwards with load_head / load_*_layer."""
        _lib.load()
        sh, dev = self.config, self.device
        self.kv_cache = self._make_kv_cache(**kwargs)
        inv_freq, scaling = llama3_inv_freq(sh)
        self.inv_freq, self.attention_scaling = inv_freq.to(dev), scaling
        cos, sin = rope_tables(inv_freq, scaling, self.max_length)
        self.cos_cache, self.sin_cache = cos.to(dev), sin.to(dev)
        H, I, Hq, Hkv, D = sh.hidden_size, sh.intermediate_size, self.num_heads, self.num_key_value_heads, self.head_dim
        f16 = dict(dtype=torch.float16, device=dev)
        self._hidden = torch.zeros(MAX_ROWS, H, **f16)
        self._normed = torch.zeros(MAX_ROWS, H, **f16)
        self._qkv = torch.zeros(MAX_ROWS, (Hq + 2 * Hkv) * D, **f16)
        self._q_rot = torch.zeros(MAX_ROWS, Hq * D, **f16)
        self._attn_out = torch.zeros(MAX_ROWS, Hq * D, **f16)
        self._act = torch.zeros(MAX_ROWS, I, **f16)
        lib = _lib.load()
        self._gemm_ws = torch.zeros(lib.umb_llama_gemm_ws_bytes(H, I, Hq, Hkv, D, sh.vocab_size), dtype=torch.uint8,
                                    device=dev)
        self._attn_ws = torch.zeros(lib.umb_tree_attn_workspace_bytes(MAX_ROWS, Hq, D, self.max_length),
                                    dtype=torch.uint8, device=dev)
        self._argmax_keys = torch.zeros(MAX_ROWS, dtype=torch.int64, device=dev)
        self.layer_weights = [None] * self.num_layers
        self._c_layers = (_lib.LlamaLayer * self.num_layers)()
        return self

    supports_graph = True        # forward_dyn (CUDA-graph replay of the whole iteration) needs everything resident

    def _make_kv_cache(self, **kwargs):
        if int(kwargs.get("kv_host_layers", 0)) > 0:
            raise RuntimeError("host-resident KV pages need the paged forward: build the model with offload=... "
                               "(models/offload.py), e.g. SyntheticLlama(..., kv_host_layers=N)")
        return KV_Cache(self.config, max_length=self.max_length, device=self.device, dtype=self.dtype,
                        batch_size=self.batch_size, num_layers=self.num_layers)

    @property
    def k_cache(self):
        return self.kv_cache.k_cache

    @property
    def v_cache(self):
        return self.kv_cache.v_cache

    @property
    def kv_offset(self):
        return self.kv_cache.kv_offset

    @kv_offset.setter
    def kv_offset(self, v):
        self.kv_cache.kv_offset = v

    # ------------------------------------------------------------------ weights
    def load_head(self, embed_tokens: torch.Tensor, lm_head: torch.Tensor | None, norm_weight: torch.Tensor):
        dev = self.device
        self.embed_tokens = embed_tokens.to(dev, torch.float16).contiguous()
        self.lm_head = self.embed_tokens if lm_head is None else lm_head.to(dev, torch.float16).contiguous()
        self.norm_weight = norm_weight.to(dev, torch.float16).contiguous()
        self._c_model = None

    def load_fp16_layer(self, idx: int, w: dict):
        """w: wq/wk/wv/wo/gate_proj/up_proj/down_proj fp16 [out, in] + the two norm vectors
        (LlamaLayer.init_parameters, llama_layer.py:27-40)."""
        dev = self.device
        I = self.config.intermediate_size
        g = lambda n: w[n].to(dev, torch.float16)
        wqkv = torch.cat([g("wq"), g("wk"), g("wv")], dim=0).contiguous()
        gate, up = g("gate_proj"), g("up_proj")
        wgu = torch.cat([gate.view(I // 64, 64, -1), up.view(I // 64, 64, -1)], dim=1).reshape(2 * I, -1).contiguous()
        self._set_layer(idx, wqkv, g("wo").contiguous(), wgu, g("down_proj").contiguous(), g("input_layernorm"),
                        g("post_attention_layernorm"))

    def _set_layer(self, idx, wqkv, wo, wgu, wdown, in_norm, post_norm):
        self.layer_weights[idx] = dict(wqkv=wqkv, wo=wo, wgu=wgu, wdown=wdown, input_norm=in_norm.contiguous(),
                                       post_norm=post_norm.contiguous())
        cl = self._c_layers[idx]
        lw = self.layer_weights[idx]
        cl.input_norm, cl.post_norm = lw["input_norm"].data_ptr(), lw["post_norm"].data_ptr()
        cl.wqkv, cl.wo, cl.wgu, cl.wdown = wqkv.data_ptr(), wo.data_ptr(), wgu.data_ptr(), wdown.data_ptr()
        self._c_model = None

    def weight_bytes(self) -> int:
        """Algorithmic weight bytes one forward reads (layers + norms + lm_head)."""
        n = self.lm_head.numel() * 2 + self.norm_weight.numel() * 2
        for lw in self.layer_weights:
            n += sum(t.numel() * t.element_size() for t in lw.values())
        return n

    def _model_struct(self) -> _lib.LlamaModel:
        if self._c_model is not None:
            return self._c_model
        assert all(lw is not None for lw in self.layer_weights), "load every layer before running the model"
        sh = self.config
        m = _lib.LlamaModel()
        m.hidden_size, m.intermediate_size, m.num_layers = sh.hidden_size, sh.intermediate_size, self.num_layers
        m.num_q_heads, m.num_kv_heads, m.head_dim, m.vocab = self.num_heads, self.num_key_value_heads, self.head_dim, sh.vocab_size
        m.rms_eps, m.awq, m.max_length = sh.rms_norm_eps, int(self.awq), self.max_length
        m.layers = C.cast(self._c_layers, C.POINTER(_lib.LlamaLayer))
        m.embed, m.final_norm, m.lm_head = self.embed_tokens.data_ptr(), self.norm_weight.data_ptr(), self.lm_head.data_ptr()
        m.cos, m.sin = self.cos_cache.data_ptr(), self.sin_cache.data_ptr()
        m.k_cache, m.v_cache = self.kv_cache.k_cache.data_ptr(), self.kv_cache.v_cache.data_ptr()
        m.hidden, m.normed, m.qkv = self._hidden.data_ptr(), self._normed.data_ptr(), self._qkv.data_ptr()
        m.q_rot, m.attn_out, m.act = self._q_rot.data_ptr(), self._attn_out.data_ptr(), self._act.data_ptr()
        m.gemm_ws, m.gemm_ws_bytes = self._gemm_ws.data_ptr(), self._gemm_ws.numel()
        m.attn_ws, m.attn_ws_bytes = self._attn_ws.data_ptr(), self._attn_ws.numel()
        self._c_model = m
        return m

    # ------------------------------------------------------------------ forward
    def _forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, mask: _lib.AttnMask, logits_from: int = 0,
                 want_logits: bool = True, argmax_keys: torch.Tensor | None = None) -> torch.Tensor | None:
        ids = input_ids.reshape(-1)
        pos = position_ids.reshape(-1)
        if not ids.is_cuda:
            raise RuntimeError("input_ids must be a CUDA tensor (there is no CPU path)")
        assert ids.dtype == torch.int64 and pos.dtype == torch.int64 and ids.is_contiguous() and pos.is_contiguous()
        rows = ids.numel()
        kv_offset = self.kv_cache.kv_offset
        if kv_offset + rows > self.max_length:
            raise RuntimeError(f"KV cache overflow: {kv_offset} + {rows} > {self.max_length}")
        logits = None
        if want_logits:
            logits = torch.empty((rows - logits_from, self.config.vocab_size), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().umb_llama_forward(
            C.byref(self._model_struct()), C.c_void_p(ids.data_ptr()), C.c_void_p(pos.data_ptr()), rows, kv_offset,
            C.byref(mask), C.c_void_p(logits.data_ptr()) if logits is not None else None, logits_from,
            C.c_void_p(argmax_keys.data_ptr()) if argmax_keys is not None else None,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "llama_forward")
        self.kv_cache.kv_offset = kv_offset + rows       # KV_Cache.update_kv_cache bumps it at layer 0 (cache.py:61-62)
        return logits

    @torch.inference_mode()
    def inference(self, input_ids: torch.LongTensor, position_ids: torch.LongTensor, attention_mask: torch.Tensor,
                  storage_ids: torch.LongTensor | None = None):
        """Reference signature (models/base.py:16-23, llama.py:305-322): fp32 logits [1, q, V].
        ``attention_mask`` is the reference's dense bool [q, >= kv]; rows are appended at kv_offset and
        the *values* of storage_ids are ignored, as KV_Cache.update_kv_cache does (cache.py:60-64)."""
        mask = attention_mask
        if mask.dtype != torch.bool and mask.dtype != torch.uint8:
            mask = mask != 0
        if mask.stride(-1) != 1:
            mask = mask.contiguous()
        rows = input_ids.numel()
        assert mask.shape[0] == rows and mask.shape[1] >= self.kv_cache.kv_offset + rows
        m = ops.make_mask(ops.UMB_MASK_DENSE, dense_mask=mask)
        logits = self._forward(input_ids, position_ids, m)
        return logits.unsqueeze(0)

    @torch.inference_mode()
    def prefill_forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, last_only: bool = True,
                        want_logits: bool = True):
        """Causal forward over new rows appended at kv_offset; logits of the last row only by default
        (the engines use nothing else, static_speculation_engine.py:172)."""
        rows = input_ids.numel()
        m = ops.make_mask(ops.UMB_MASK_CAUSAL, prefix_len=self.kv_cache.kv_offset)
        return self._forward(input_ids, position_ids, m, logits_from=rows - 1 if last_only else 0,
                             want_logits=want_logits)

    @torch.inference_mode()
    def verify_forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, prefix_len: int,
                       tree_bits: torch.Tensor, argmax_keys: torch.Tensor | None = None, want_logits: bool = True):
        """Fast path of verify() and of the draft's per-level forward: the mask is (prefix_len, ancestor
        bitmask rows), never materialised."""
        m = ops.make_mask(ops.UMB_MASK_TREE, prefix_len=prefix_len, tree_bits=tree_bits)
        return self._forward(input_ids, position_ids, m, argmax_keys=argmax_keys, want_logits=want_logits)

    @torch.inference_mode()
    def forward_dyn(self, tokens_base: torch.Tensor, position_base: torch.Tensor, rows: int, n_dev: torch.Tensor,
                    row_offset: int, kv_len_bound: int, tree_bits: torch.Tensor, logits: torch.Tensor | None = None,
                    argmax_keys: torch.Tensor | None = None):
        """Graph-replay form of verify_forward: the committed prefix length is read on the device from
        ``n_dev`` (int32), so a captured CUDA graph of this call follows the moving prefix.  Host-side
        ``kv_offset`` bookkeeping is the caller's job."""
        assert tokens_base.dtype == torch.int64 and position_base.dtype == torch.int64 and n_dev.dtype == torch.int32
        _lib.check(_lib.load().umb_llama_forward_dyn(
            C.byref(self._model_struct()), C.c_void_p(tokens_base.data_ptr()), C.c_void_p(position_base.data_ptr()), rows,
            C.c_void_p(n_dev.data_ptr()), row_offset, kv_len_bound, C.c_void_p(tree_bits.data_ptr()), tree_bits.shape[1],
            C.c_void_p(logits.data_ptr()) if logits is not None else None,
            C.c_void_p(argmax_keys.data_ptr()) if argmax_keys is not None else None,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "llama_forward_dyn")
        return logits

    # ------------------------------------------------------------------ graph surface (base.py:25-31, llama.py:506-533)
    storage_id_semantics = False        # True for a model made with ``cuda_graph=True`` (the reference's StaticKV_Cache)

    def initialize_cuda_graph(self, decoding_seqlens, n_warmups: int = 12):
        """``LlamaCudagraph.initialize_cuda_graph`` captures one torch graph per decoding length so that a forward
        stops being ~30 eager launches per layer.  Here a forward of any length is one C call that enqueues every
        launch, and the engines capture whole iterations, so there is nothing to capture ahead of time: the lengths
        are recorded and, as in the reference (llama.py:519), the cache is cleared."""
        self.graph_lengths = sorted(set(getattr(self, "graph_lengths", [])) | {int(x) for x in decoding_seqlens})
        self.clear()

    @torch.inference_mode()
    def graph_inference(self, input_ids: torch.LongTensor, *, storage_ids: torch.LongTensor | None = None,
                        position_ids: torch.LongTensor | None = None, attention_mask: torch.Tensor | None = None):
        """The call the reference's examples make on every model (examples/bench.py:35-47, generate.py:55-86,
        construct_sequoia.py:71-75), keyword arguments only (its two definitions order them differently).
        Plain models: ``inference`` (base.py:25-31).  ``cuda_graph=True`` models: the new rows go where
        ``storage_ids`` says (StaticKV_Cache.update_kv_cache, cache.py:150-158) -- a contiguous ascending range, which
        becomes the append position; the attention then spans the rows up to its end, as the mask beyond is empty."""
        if self.storage_id_semantics and storage_ids is not None and storage_ids.numel() > 0:
            first, last = storage_ids.reshape(-1)[[0, -1]].tolist()                 # one host read
            if last - first != storage_ids.numel() - 1:
                raise ValueError("storage_ids must be a contiguous ascending range")
            self.kv_cache.kv_offset = int(first)
        return self.inference(input_ids, position_ids, attention_mask, storage_ids)

    # ------------------------------------------------------------------ cache ops (llama.py:136-142)
    def gather_kv_incremental(self, indices, offset: int):
        self.kv_cache.gather_kv_incremental(indices, offset)

    def clear(self):
        self.kv_cache.clear()


class LlamaAwq(Llama):
    """AWQ INT4 Llama (reference ``LlamaAwq``, llama.py:222-322): the seven linears of a layer become
    three fused packed streams (q|k|v, gate|up interleaved for the SiLU epilogue) + o + down."""
    awq = True

    def load_awq_layer(self, idx: int, w: dict):
        """w[name] = (qweight int32 [K, N/8], qzeros int32 [K/g, N/8], scales fp16 [K/g, N]) for the
        seven linears (LlamaAwqLayer.init_parameters, llama_layer.py:176-218) + the two norm vectors."""
        dev, sh = self.device, self.config
        H, I, D = sh.hidden_size, sh.intermediate_size, self.head_dim
        nq, nkv = self.num_heads * D, self.num_key_value_heads * D
        d = lambda n: [t.to(dev) for t in w[n]]
        qkv_n = nq + 2 * nkv
        wqkv = ops.awq_repack(*d("wq"), N_total=qkv_n)
        ops.awq_repack(*d("wk"), packed=wqkv, N_total=qkv_n, row_offset=nq)
        ops.awq_repack(*d("wv"), packed=wqkv, N_total=qkv_n, row_offset=nq + nkv)
        wo = ops.awq_repack(*d("wo"))
        wgu = ops.awq_repack(*d("gate_proj"), N_total=2 * I, row_block=64, row_stride=128)
        ops.awq_repack(*d("up_proj"), packed=wgu, N_total=2 * I, row_block=64, row_stride=128, row_offset=64)
        wdown = ops.awq_repack(*d("down_proj"))
        self._set_layer(idx, wqkv, wo, wgu, wdown, w["input_layernorm"].to(dev, torch.float16),
                        w["post_attention_layernorm"].to(dev, torch.float16))


# --------------------------------------------------------------------------------------------
# synthetic weights: seeded, real shapes; value-independent timing
# --------------------------------------------------------------------------------------------
LINEAR_NAMES = ("wq", "wk", "wv", "wo", "gate_proj", "up_proj", "down_proj")


def linear_dims(shape: LlamaShape) -> dict[str, tuple[int, int]]:
    h, i = shape.hidden_size, shape.intermediate_size
    kvd = shape.num_key_value_heads * shape.head_dim
    return {"wq": (h, h), "wk": (h, kvd), "wv": (h, kvd), "wo": (h, h),
            "gate_proj": (h, i), "up_proj": (h, i), "down_proj": (i, h)}


def _synth_awq_linear(k, n, group, gen, scale=0.003):
    dev = gen.device
    qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), generator=gen, dtype=torch.int64, device=dev).to(torch.int32)
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (k // group, n // 8), generator=gen, dtype=torch.int64, device=dev).to(torch.int32)
    scales = (scale * (0.5 + torch.rand((k // group, n), generator=gen, device=dev))).to(torch.float16)
    return qweight, qzeros, scales


def synth_layer(shape: LlamaShape, layer_idx: int, awq: bool, seed: int, gen_device="cpu") -> dict:
    """Same draw order as the test oracle's generator, so a CPU-generated model is weight-identical
    to it for a given seed."""
    gen = torch.Generator(device=gen_device).manual_seed(seed + layer_idx)
    out = {}
    for name, (k, n) in linear_dims(shape).items():
        if awq:
            out[name] = _synth_awq_linear(k, n, shape.group_size, gen)
        else:
            out[name] = (0.02 * torch.randn((n, k), generator=gen, device=gen_device)).to(torch.float16)
    out["input_layernorm"] = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen, device=gen_device)).to(torch.float16)
    out["post_attention_layernorm"] = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen, device=gen_device)).to(torch.float16)
    return out


def synth_head(shape: LlamaShape, seed: int, gen_device="cpu") -> dict:
    gen = torch.Generator(device=gen_device).manual_seed(seed - 1)
    embed = (0.02 * torch.randn((shape.vocab_size, shape.hidden_size), generator=gen, device=gen_device)).to(torch.float16)
    lm_head = None if shape.tie_word_embeddings else \
        (0.02 * torch.randn((shape.vocab_size, shape.hidden_size), generator=gen, device=gen_device)).to(torch.float16)
    norm = (1 + 0.1 * torch.randn(shape.hidden_size, generator=gen, device=gen_device)).to(torch.float16)
    return {"embed_tokens": embed, "lm_head": lm_head, "norm": norm}


def SyntheticLlama(shape: LlamaShape, max_length: int, awq: bool, seed: int = 1234, device: str = "cuda:0",
                   num_layers: int | None = None, gen_device: str = "cpu", offload: dict | None = None,
                   kv_host_layers: int = 0):
    """A ready-to-run model with seeded synthetic weights of the given architecture.  ``offload`` (e.g.
    ``{"num_cache_layers": 0}``) builds the paged variant of models/offload.py with those alloc kwargs;
    ``kv_host_layers`` keeps the K/V planes of that many upper layers in pinned host memory (models/kv_paging.py),
    weights resident unless ``offload`` says otherwise."""
    if kv_host_layers:
        n = shape.num_hidden_layers if num_layers is None else num_layers
        offload = dict({"num_cache_layers": n}, **(offload or {}), kv_host_layers=kv_host_layers)
    if offload is not None:
        from .offload import LlamaAwqOffload, LlamaOffload
        model = (LlamaAwqOffload if awq else LlamaOffload)(shape, max_length=max_length, device=device,
                                                           num_layers=num_layers)
        model.alloc(**offload)
    else:
        model = (LlamaAwq if awq else Llama)(shape, max_length=max_length, device=device, num_layers=num_layers)
        model.alloc()
    head = synth_head(shape, seed, gen_device)
    model.load_head(head["embed_tokens"], head["lm_head"], head["norm"])
    for i in range(model.num_layers):
        w = synth_layer(shape, i, awq, seed, gen_device)
        (model.load_awq_layer if awq else model.load_fp16_layer)(i, w)
        del w
    torch.cuda.synchronize(model.device)
    return model

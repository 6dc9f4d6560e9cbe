"""Tensor-level wrappers over the C ABI (torch tensors in, raw pointers out).

PyTorch is plumbing here: device memory, the current stream, dtype/shape checks.  Every function
enqueues hand-written sm_100a kernels on ``torch.cuda.current_stream()`` and returns immediately.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (AttnMask, Epilogue, RopeEpilogue, UMB_EPI_LOGITS, UMB_EPI_RESIDUAL, UMB_EPI_ROPE_QKV, UMB_EPI_SILU_MUL,
                   UMB_EPI_STORE, UMB_MASK_CAUSAL, UMB_MASK_DENSE, UMB_MASK_TREE, UMB_SAMPLE_ARGMAX, UMB_SAMPLE_DYNAMIC,
                   UMB_SAMPLE_STATIC, check)

GROUP = 128
SAMPLE_MAX_TOPK = 1024      # umb_sample_f32 keeps the top-k survivors of a row in shared memory


def _p(t: torch.Tensor | None):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda(t: torch.Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


# ---------------------------------------------------------------------------------------------
# (i) AWQ
# ---------------------------------------------------------------------------------------------
def awq_packed_bytes(K: int, N_total: int, group: int = GROUP) -> int:
    n = _lib.load().umb_awq_packed_bytes(K, N_total, group)
    if n == 0:
        raise ValueError(f"unsupported AWQ shape K={K} N={N_total} group={group}")
    return n


def awq_repack(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, group: int = GROUP,
               packed: torch.Tensor | None = None, N_total: int | None = None, row_block: int | None = None,
               row_stride: int | None = None, row_offset: int = 0) -> torch.Tensor:
    """AutoAWQ GEMM-layout triplet -> (rows of) the K-major packed stream.  Several matrices sharing K
    can be packed into one stream (q|k|v, gate|up) via (row_block, row_stride, row_offset)."""
    _cuda(qweight, torch.int32, "qweight"), _cuda(qzeros, torch.int32, "qzeros"), _cuda(scales, torch.float16, "scales")
    K, N = qweight.shape[0], qweight.shape[1] * 8
    N_total = N if N_total is None else N_total
    row_block = N if row_block is None else row_block
    row_stride = row_block if row_stride is None else row_stride
    if packed is None:
        packed = torch.zeros(awq_packed_bytes(K, N_total, group), dtype=torch.uint8, device=qweight.device)
    check(_lib.load().umb_awq_repack(_p(qweight.contiguous()), _p(qzeros.contiguous()), _p(scales.contiguous()), K, N,
                                     group, _p(packed), N_total, row_block, row_stride, row_offset, _stream()),
          "awq_repack")
    return packed


def awq_dequantize(qweight, qzeros, scales, group: int = GROUP) -> torch.Tensor:
    """awq_ext.dequantize_weights_cuda drop-in: fp16 [K, N]."""
    K, N = qweight.shape[0], qweight.shape[1] * 8
    out = torch.empty((K, N), dtype=torch.float16, device=qweight.device)
    check(_lib.load().umb_awq_dequantize_f16(_p(qweight.contiguous()), _p(qzeros.contiguous()), _p(scales.contiguous()),
                                             K, N, group, _p(out), _stream()), "awq_dequantize")
    return out


def awq_unpack(packed: torch.Tensor, K: int, N_total: int, group: int = GROUP) -> torch.Tensor:
    out = torch.empty((N_total, K), dtype=torch.float16, device=packed.device)
    check(_lib.load().umb_awq_unpack_f16(_p(packed), K, N_total, group, _p(out), _stream()), "awq_unpack")
    return out


def gemm_workspace(M: int, K: int, N_total: int, device) -> torch.Tensor:
    n = _lib.load().umb_gemm_workspace_bytes(M, K, N_total)
    return torch.zeros(n, dtype=torch.uint8, device=device)


def make_rope_epilogue(q_out, k_cache, v_cache, cos, sin, position_ids, Hq: int, Hkv: int, D: int, kv_offset: int,
                       layout: str = "NHD", n_dev=None, row_offset: int = 0) -> RopeEpilogue:
    """Operands of UMB_EPI_ROPE_QKV: k/v = one layer of the cache ([S, Hkv, D] "NHD" or [Hkv, S, D] "HND")."""
    r = RopeEpilogue()
    r.q_out, r.k_cache, r.v_cache = q_out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    r.cos, r.sin, r.position_ids = cos.data_ptr(), sin.data_ptr(), position_ids.data_ptr()
    r.n_dev = n_dev.data_ptr() if n_dev is not None else None
    r.row_offset, r.kv_offset = row_offset, kv_offset
    r.num_q_heads, r.num_kv_heads, r.head_dim = Hq, Hkv, D
    r.kv_tok_stride, r.kv_head_stride = (k_cache.stride(0), k_cache.stride(1)) if layout == "NHD" else \
        (k_cache.stride(1), k_cache.stride(0))
    return r


def _epilogue(mode, bias=None, residual=None, argmax=None, rope=None):
    e = Epilogue()
    e.mode = mode
    e.bias = bias.data_ptr() if bias is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.ld_residual = residual.stride(0) if residual is not None else 0
    e.argmax = argmax.data_ptr() if argmax is not None else None
    if rope is not None:
        e._rope_keepalive = rope                       # the struct is read at launch time
        e.rope = C.addressof(rope)
    return e


def _gemm_out(x, N_out, mode, out):
    M = x.shape[0]
    if out is None:
        out = torch.empty((M, N_out), dtype=torch.float32 if mode == UMB_EPI_LOGITS else torch.float16, device=x.device)
    return out


def awq_gemm(x: torch.Tensor, packed: torch.Tensor, K: int, N_total: int, *, mode: int = UMB_EPI_STORE, bias=None,
             residual=None, argmax=None, out: torch.Tensor | None = None, workspace: torch.Tensor | None = None,
             group: int = GROUP, rope: RopeEpilogue | None = None) -> torch.Tensor:
    """y[M, N] = x[M, K] . W from the packed stream (M <= 128)."""
    _cuda(x, torch.float16, "x")
    assert x.dim() == 2 and x.shape[1] == K and x.stride(1) == 1
    N_out = N_total // 2 if mode == UMB_EPI_SILU_MUL else N_total
    out = _gemm_out(x, N_out, mode, out)
    if workspace is None:
        workspace = gemm_workspace(x.shape[0], K, N_total, x.device)
    e = _epilogue(mode, bias, residual, argmax, rope)
    check(_lib.load().umb_awq_gemm_f16(_p(x), x.stride(0), _p(packed), _p(out), out.stride(0), x.shape[0], K, N_total,
                                       group, C.byref(e), _p(workspace), workspace.numel(), _stream()), "awq_gemm")
    return out


def awq_gemm_simt(x, packed, K, N_total):
    out = torch.empty((x.shape[0], N_total), dtype=torch.float16, device=x.device)
    check(_lib.load().umb_awq_gemm_f16_simt(_p(x), x.stride(0), _p(packed), _p(out), out.stride(0), x.shape[0], K,
                                            N_total, _stream()), "awq_gemm_simt")
    return out


def gemm_f16(x: torch.Tensor, w: torch.Tensor, *, mode: int = UMB_EPI_STORE, bias=None, residual=None, argmax=None,
             out=None, workspace=None, rope: RopeEpilogue | None = None) -> torch.Tensor:
    """y = x . w^T for fp16 w [N, K] (torch Linear layout), M <= 128."""
    _cuda(x, torch.float16, "x"), _cuda(w, torch.float16, "w")
    N, K = w.shape
    assert x.shape[1] == K and x.stride(1) == 1 and w.stride(1) == 1
    N_out = N // 2 if mode == UMB_EPI_SILU_MUL else N
    out = _gemm_out(x, N_out, mode, out)
    if workspace is None:
        workspace = gemm_workspace(x.shape[0], K, N, x.device)
    e = _epilogue(mode, bias, residual, argmax, rope)
    check(_lib.load().umb_gemm_f16(_p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), x.shape[0], K, N,
                                   C.byref(e), _p(workspace), workspace.numel(), _stream()), "gemm_f16")
    return out


def probe_umma(a: torch.Tensor, b: torch.Tensor, mode: int) -> torch.Tensor:
    d = torch.empty((128, 128 if mode >= 2 else 16), dtype=torch.float32, device=a.device)
    check(_lib.load().umb_probe_umma(_p(a), _p(b), _p(d), mode, _stream()), "probe_umma")
    return d


# ---------------------------------------------------------------------------------------------
# (ii) attention
# ---------------------------------------------------------------------------------------------
def make_mask(mode: int, prefix_len: int = 0, tree_bits: torch.Tensor | None = None,
              dense_mask: torch.Tensor | None = None, dyn: torch.Tensor | None = None) -> AttnMask:
    m = AttnMask()
    m.mask_mode = mode
    m.prefix_len = prefix_len
    m.dyn = dyn.data_ptr() if dyn is not None else None
    if tree_bits is not None:
        _cuda(tree_bits, None, "tree_bits")
        assert tree_bits.dtype in (torch.int32, torch.uint32) and tree_bits.is_contiguous()
        m.tree_bits = tree_bits.data_ptr()
        m.tree_words = tree_bits.shape[1]
    if dense_mask is not None:
        _cuda(dense_mask, None, "dense_mask")
        assert dense_mask.dtype in (torch.bool, torch.uint8) and dense_mask.stride(1) == 1
        m.dense_mask = dense_mask.data_ptr()
        m.mask_stride = dense_mask.stride(0)
    return m


def tree_attn_workspace(qo: int, Hq: int, D: int, device) -> torch.Tensor:
    return torch.zeros(_lib.load().umb_tree_attn_workspace_bytes(qo, Hq, D, 0), dtype=torch.uint8, device=device)


def tree_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, mask: AttnMask,
              layout: str = "NHD", sm_scale: float = 0.0, soft_cap: float = 0.0, out=None, workspace=None):
    """q [qo, Hq, D]; k/v one layer of the cache: [S, Hkv, D] ("NHD") or [Hkv, S, D] ("HND")."""
    _cuda(q, torch.float16, "q")
    qo, Hq, D = q.shape
    if layout == "NHD":
        Hkv, tok, head = k_cache.shape[1], k_cache.stride(0), k_cache.stride(1)
    else:
        Hkv, tok, head = k_cache.shape[0], k_cache.stride(1), k_cache.stride(0)
    if out is None:
        out = torch.empty_like(q)
    if workspace is None:
        workspace = tree_attn_workspace(qo, Hq, D, q.device)
    check(_lib.load().umb_tree_attn_f16(_p(q), _p(k_cache), _p(v_cache), _p(out), qo, kv_len, Hq, Hkv, D, tok, head,
                                        C.byref(mask), sm_scale, soft_cap, _p(workspace), workspace.numel(), _stream()),
          "tree_attn")
    return out


def rope_append(qkv: torch.Tensor, q_out: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, cos, sin,
                position_ids: torch.Tensor, Hq: int, Hkv: int, D: int, kv_offset: int, layout: str = "NHD"):
    qo = qkv.shape[0]
    tok, head = (k_cache.stride(0), k_cache.stride(1)) if layout == "NHD" else (k_cache.stride(1), k_cache.stride(0))
    check(_lib.load().umb_rope_append_f16(_p(qkv), qkv.stride(0), _p(q_out), _p(k_cache), _p(v_cache), _p(cos), _p(sin),
                                          _p(position_ids), qo, Hq, Hkv, D, tok, head, kv_offset, None, _stream()),
          "rope_append")


# ---------------------------------------------------------------------------------------------
# (iii) accept / gather
# ---------------------------------------------------------------------------------------------
def tree_accept(sampled, tokens, position_ids, parents, depth, num_nodes: int, eos: torch.Tensor | None,
                result: torch.Tensor, state: torch.Tensor | None = None):
    N = parents.numel()
    _cuda(sampled, torch.int64, "sampled"), _cuda(tokens, torch.int64, "tokens")
    _cuda(position_ids, torch.int64, "position_ids"), _cuda(parents, torch.int32, "parents")
    _cuda(depth, torch.int32, "depth"), _cuda(result, torch.int32, "result")
    assert result.numel() >= 4 + N
    check(_lib.load().umb_tree_accept(_p(sampled), _p(tokens), _p(position_ids), _p(parents), _p(depth), N, num_nodes,
                                      min(tokens.numel(), position_ids.numel()), _p(state), _p(eos), 0 if eos is None else eos.numel(), _p(result), _stream()),
          "tree_accept")


def store_tokens(tokens: torch.Tensor, n_dev: torch.Tensor, offset: int, src: torch.Tensor):
    """tokens[n_dev[0] + offset + i] = src[i] with n read on the device (graph replay)."""
    _cuda(tokens, torch.int64, "tokens"), _cuda(src, torch.int64, "src"), _cuda(n_dev, torch.int32, "n_dev")
    check(_lib.load().umb_store_tokens(_p(tokens), _p(n_dev), offset, _p(src), src.numel(), tokens.numel(), _stream()),
          "store_tokens")


TOPK_MAX = 64               # umb_topk_f32


def topk(logits: torch.Tensor, k: int, out_vals: torch.Tensor | None = None, out_ids: torch.Tensor | None = None):
    """Row-wise top-k (values descending, lower id first among equals) in ONE launch -> (vals fp32, ids int64) [rows, k]."""
    _cuda(logits, torch.float32, "logits")
    assert logits.dim() == 2 and logits.stride(1) == 1
    rows, vocab = logits.shape
    if out_vals is None:
        out_vals = torch.empty((rows, k), dtype=torch.float32, device=logits.device)
    if out_ids is None:
        out_ids = torch.empty((rows, k), dtype=torch.int64, device=logits.device)
    check(_lib.load().umb_topk_f32(_p(logits), logits.stride(0), rows, vocab, k, _p(out_vals), _p(out_ids), _stream()), "topk")
    return out_vals, out_ids


def store_tokens_gather(tokens: torch.Tensor, src: torch.Tensor, gather: torch.Tensor, offset: int, *, num_nodes: int = 0,
                        n_dev: torch.Tensor | None = None):
    """tokens[n + offset + i] = src.flatten()[gather[i]], n on the host (num_nodes) or on the device (n_dev)."""
    _cuda(tokens, torch.int64, "tokens"), _cuda(src, torch.int64, "src"), _cuda(gather, torch.int64, "gather")
    assert src.is_contiguous() and gather.is_contiguous()
    check(_lib.load().umb_store_tokens_gather(_p(tokens), _p(n_dev), num_nodes, offset, _p(src), _p(gather), gather.numel(),
                                              tokens.numel(), _stream()), "store_tokens_gather")


def dyn_tree_expand(topk_logits: torch.Tensor, topk_ids: torch.Tensor, width: int, level_first: int, num_nodes: int,
                    tree_score: torch.Tensor, tokens: torch.Tensor, parents: torch.Tensor, tree_bits: torch.Tensor,
                    n_dev: torch.Tensor | None = None):
    """One level of the dynamic tree builder (dynamic_speculation_engine.py:231-248): beam scores ->
    the ``width`` best children, their tokens, parents, scores and ancestor-bit rows."""
    _cuda(topk_logits, torch.float32, "topk_logits"), _cuda(topk_ids, torch.int64, "topk_ids")
    _cuda(tree_score, torch.float32, "tree_score"), _cuda(tokens, torch.int64, "tokens")
    _cuda(parents, torch.int32, "parents")
    assert tree_bits.dtype in (torch.int32, torch.uint32) and tree_bits.is_contiguous()
    assert topk_logits.is_contiguous() and topk_ids.is_contiguous() and topk_logits.shape == topk_ids.shape
    rows, beams = topk_logits.shape
    check(_lib.load().umb_dyn_tree_expand(_p(topk_logits), _p(topk_ids), rows, beams, width, level_first, num_nodes,
                                          tokens.numel(), _p(n_dev), _p(tree_score), _p(tokens), _p(parents), _p(tree_bits),
                                          tree_bits.shape[1], _stream()), "dyn_tree_expand")


def kv_gather(k_cache: torch.Tensor, v_cache: torch.Tensor, result: torch.Tensor, offset: int, layout: str = "LSHD",
              zero_upto: int = 0, offset_dev: torch.Tensor | None = None):
    """In-place compaction of accepted rows for every layer.  k/v: [L, S, Hkv, D] ("LSHD", KV_Cache)
    or [L, Hkv, S, D] ("LHSD", StaticKV_Cache)."""
    es = k_cache.element_size()
    if layout == "LSHD":
        L, S, Hkv, D = k_cache.shape
        planes, plane_stride, row_stride, row_bytes = L, k_cache.stride(0) * es, k_cache.stride(1) * es, Hkv * D * es
    else:
        L, Hkv, S, D = k_cache.shape
        planes, plane_stride, row_stride, row_bytes = L * Hkv, k_cache.stride(1) * es, k_cache.stride(2) * es, D * es
    check(_lib.load().umb_kv_gather(_p(k_cache), _p(v_cache), planes, plane_stride, row_stride, row_bytes, _p(result),
                                    offset, _p(offset_dev) if offset_dev is not None else None, zero_upto, _stream()),
          "kv_gather")


# ---------------------------------------------------------------------------------------------
# glue
# ---------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out=None):
    x2 = x.reshape(-1, x.shape[-1])
    out = torch.empty_like(x2) if out is None else out
    check(_lib.load().umb_rmsnorm_f16(_p(x2), _p(w), _p(out), x2.shape[0], x2.shape[1], eps, _stream()), "rmsnorm")
    return out.reshape(x.shape)


def silu_mul(gate, up):
    out = torch.empty_like(gate)
    check(_lib.load().umb_silu_mul_f16(_p(gate), _p(up), _p(out), gate.numel(), _stream()), "silu_mul")
    return out


def add(a, b):
    out = torch.empty_like(a)
    check(_lib.load().umb_add_f16(_p(a), _p(b), _p(out), a.numel(), _stream()), "add")
    return out


def embedding(ids: torch.Tensor, table: torch.Tensor):
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float16, device=table.device)
    check(_lib.load().umb_embedding_f16(_p(ids), _p(table), _p(out), ids.numel(), table.shape[1], _stream()), "embedding")
    return out


def argmax(logits: torch.Tensor) -> torch.Tensor:
    _cuda(logits, torch.float32, "logits")
    rows, vocab = logits.shape
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    check(_lib.load().umb_argmax_f32(_p(logits), logits.stride(0), rows, vocab, _p(out), _stream()), "argmax")
    return out


def sample(logits: torch.Tensor, mode: int, *, temperature: float = 1.0, top_k: int = 32, top_p: float = 0.9,
           uniforms: torch.Tensor | None = None, penalty_ids: torch.Tensor | None = None, penalty: float = 1.0,
           penalty_count: int | None = None, penalty_count_dev: torch.Tensor | None = None,
           out: torch.Tensor | None = None) -> torch.Tensor:
    """Fused target sampling (umb_sample_f32): repetition penalty -> top-k -> softmax(T) -> top-p -> one token per
    row.  ``uniforms``: fp32 [rounds, rows] (UMB_SAMPLE_STATIC, the engine's fixed uniform_samples) or [rows]
    (UMB_SAMPLE_DYNAMIC).  ``penalty_ids``: int64 ids seen so far (first ``penalty_count`` entries, plus
    ``penalty_count_dev[0]`` when given)."""
    _cuda(logits, torch.float32, "logits")
    assert logits.dim() == 2 and logits.stride(1) == 1
    rows, vocab = logits.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    rounds = 0
    if mode != UMB_SAMPLE_ARGMAX:
        _cuda(uniforms, torch.float32, "uniforms")
        assert uniforms.is_contiguous() and uniforms.shape[-1] == rows
        rounds = uniforms.shape[0] if uniforms.dim() == 2 else 1
    count = 0
    if penalty_ids is not None:
        _cuda(penalty_ids, torch.int64, "penalty_ids")
        assert penalty_ids.is_contiguous()
        count = penalty_ids.numel() if penalty_count is None else penalty_count
    check(_lib.load().umb_sample_f32(_p(logits), logits.stride(0), rows, vocab, mode, float(temperature), int(top_k),
                                     float(top_p), _p(uniforms), rounds, _p(penalty_ids), count, _p(penalty_count_dev),
                                     float(penalty), _p(out), _stream()), "sample")
    return out


def argmax_keys_to_ids(keys: torch.Tensor) -> torch.Tensor:
    out = torch.empty(keys.numel(), dtype=torch.int64, device=keys.device)
    check(_lib.load().umb_argmax_keys_to_ids(_p(keys), keys.numel(), _p(out), _stream()), "argmax_keys_to_ids")
    return out

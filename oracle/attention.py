"""Oracle: tree-masked grouped-query attention (numpy, float64 maths on fp16 inputs).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates ``KV_Cache.compute_attention`` (umbrella/attn/cache.py:67-87), i.e.
``flashinfer.single_prefill_with_kv_cache(q[qo,Hq,D], k[kv,Hkv,D], v[kv,Hkv,D], kv_layout="NHD",
custom_mask=bool[qo,kv])`` with the default ``sm_scale = 1/sqrt(D)``, and its pure-torch twin
``StaticKV_Cache.compute_attention`` (cache.py:169-192), which is runnable on CPU and is what
``oracle/make_golden.py`` pins this file against.
"""
from __future__ import annotations

import math

import numpy as np


def kv_append(k_cache: np.ndarray, v_cache: np.ndarray, new_k: np.ndarray, new_v: np.ndarray,
              layer: int, kv_offset: int) -> int:
    """``KV_Cache.update_kv_cache`` (cache.py:53-65): append at [kv_offset, kv_offset+m) of layer
    ``layer``; the caller bumps kv_offset once per forward (the reference does it at layer 0).
    Cache layout [L, S, Hkv, D]."""
    m = new_k.shape[0]
    k_cache[layer, kv_offset:kv_offset + m] = new_k
    v_cache[layer, kv_offset:kv_offset + m] = new_v
    return kv_offset + m


def masked_gqa_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray,
                         sm_scale: float | None = None, soft_cap: float = 0.0) -> np.ndarray:
    """q [qo, Hq, D], k/v [kv, Hkv, D], mask bool [qo, kv] (True = attend) -> float64 [qo, Hq, D].

    Query head h reads kv head h // (Hq // Hkv) (cache.py:176: the q heads of one group are
    contiguous).  Softmax in float64 over the masked row.  ``soft_cap`` > 0 applies
    cap * tanh(s / cap) to the scaled scores (Gemma only, cache.py:84).
    """
    qo, hq, d = q.shape
    kv, hkv, _ = k.shape
    g = hq // hkv
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(d)
    q64 = q.astype(np.float64)
    k64 = k.astype(np.float64)
    v64 = v.astype(np.float64)
    out = np.zeros((qo, hq, d), dtype=np.float64)
    m = np.asarray(mask, dtype=bool)[:, :kv]
    for h in range(hq):
        s = q64[:, h, :] @ k64[:, h // g, :].T * sm_scale            # [qo, kv]
        if soft_cap > 0:
            s = soft_cap * np.tanh(s / soft_cap)
        s = np.where(m, s, -np.inf)
        s = s - s.max(axis=-1, keepdims=True)
        p = np.exp(s)
        p = p / p.sum(axis=-1, keepdims=True)
        out[:, h, :] = p @ v64[:, h // g, :]
    return out


def tree_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, prefix_len: int,
                   tree_mask: np.ndarray, sm_scale: float | None = None) -> np.ndarray:
    """Same as masked_gqa_attention with the mask given in the product's compact form: row i sees
    [0, prefix_len) and prefix_len + j for every j with tree_mask[i, j]."""
    from .tree import dense_rows_from_compact
    kv = k.shape[0]
    return masked_gqa_attention(q, k, v, dense_rows_from_compact(prefix_len, np.asarray(tree_mask, bool), kv),
                                sm_scale)

"""Oracle: AWQ INT4 (GEMM layout) pack / unpack / dequantise / GEMM in numpy.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference call site is ``AwqLinear.apply`` (umbrella/quantization/awq_utils.py:63-86):

    out = awq_ext.gemm_forward_cuda(x2d, qweight, scales, qzeros, 8)          # x rows < 1024
    out = matmul(x, awq_ext.dequantize_weights_cuda(qweight, scales, qzeros, 0, 0, 0, False))

``awq_ext`` is the un-vendored dependency autoawq-kernels==0.0.8 (requirements.txt:7); the tensor
layout is defined by autoawq==0.2.7.post3 ``WQLinear_GEMM`` (requirements.txt:6).  Neither is in
/root/reference, and the reference holds no golden vector for them -> the *arithmetic* below is
"parity unpinned": it restates the published algorithm

    W[k, 8c + j] = (nib(qweight[k, c], p_j) - nib(qzeros[k // g, c], p_j)) * scales[k // g, 8c + j]
    p_j = [0, 4, 1, 5, 2, 6, 3, 7][j]            (nibble p = bits 4p .. 4p+3)
    y = x @ W

The *format* (nibble order) is pinned against the in-image vLLM restatement of the AutoAWQ packer
(vllm/model_executor/layers/quantization/utils/quant_utils.py:794-814 ``awq_pack``), see
``oracle/make_golden.py`` and ``tests/test_oracle_awq.py``.

Rounding model of the upstream kernel (for the "reference-faithful" emulation only):
  * dequantised weight is rounded to fp16:  w16 = fp16((q - z) * s)   ((q-z) exact, one rounding)
  * products accumulate in fp32 inside mma.sync; each of the 8 split-K partial sums is rounded
    to fp16, and the 8 partials are summed (torch ``.sum(0)`` on fp16: fp32 accumulate, one
    final rounding).
"""
from __future__ import annotations

import numpy as np

PACK = 8
# column j of an 8-column block lives in nibble AWQ_NIBBLE_OF_COL[j]
AWQ_NIBBLE_OF_COL = np.array([0, 4, 1, 5, 2, 6, 3, 7])
# nibble i holds column AWQ_COL_OF_NIBBLE[i]  (the packer's "interleave" [0,2,4,6,1,3,5,7])
AWQ_COL_OF_NIBBLE = np.array([0, 2, 4, 6, 1, 3, 5, 7])


def pack_cols_awq(q: np.ndarray) -> np.ndarray:
    """[R, C] integers in [0, 15] -> int32 [R, C/8] in AWQ-GEMM nibble order."""
    q = np.asarray(q)
    r, c = q.shape
    assert c % PACK == 0
    qb = q.reshape(r, c // PACK, PACK).astype(np.uint32)
    out = np.zeros((r, c // PACK), dtype=np.uint32)
    for j in range(PACK):
        out |= (qb[:, :, j] & np.uint32(0xF)) << np.uint32(4 * AWQ_NIBBLE_OF_COL[j])
    return out.view(np.int32)


def unpack_cols_awq(packed: np.ndarray) -> np.ndarray:
    """int32 [R, C/8] -> uint8 [R, C] (values 0..15), inverse of pack_cols_awq."""
    p = np.ascontiguousarray(packed).view(np.uint32)
    r, cb = p.shape
    out = np.zeros((r, cb, PACK), dtype=np.uint8)
    for j in range(PACK):
        out[:, :, j] = (p >> np.uint32(4 * AWQ_NIBBLE_OF_COL[j])) & np.uint32(0xF)
    return out.reshape(r, cb * PACK)


def dequantize(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray, group: int,
               dtype=np.float64) -> np.ndarray:
    """Exact (q - z) * s in ``dtype`` -- [K, N].  With dtype=np.float16 this is the upstream
    ``dequantize_weights_cuda`` result (one rounding of the exact product)."""
    q = unpack_cols_awq(qweight).astype(np.int32)          # [K, N]
    z = unpack_cols_awq(qzeros).astype(np.int32)           # [K/g, N]
    k = q.shape[0]
    zi = np.repeat(z, group, axis=0)[:k]
    si = np.repeat(np.asarray(scales, dtype=np.float16).astype(np.float64), group, axis=0)[:k]
    w = (q - zi).astype(np.float64) * si                   # exact in f64 (5-bit int x 11-bit mantissa)
    return w.astype(dtype)


def gemm_exact(x: np.ndarray, qweight, qzeros, scales, group: int) -> np.ndarray:
    """y = x @ W in float64 with W exact -- the ground truth every implementation is measured
    against (fp16 inputs, no intermediate rounding)."""
    w = dequantize(qweight, qzeros, scales, group, np.float64)
    return np.asarray(x, dtype=np.float16).astype(np.float64) @ w


def gemm_fp16_weights(x: np.ndarray, qweight, qzeros, scales, group: int) -> np.ndarray:
    """y = x @ fp16(W), accumulated in float64, rounded once to fp16.  This is what the product's
    tcgen05 kernel computes up to fp32 accumulation order (it rounds (q-z)*s to fp16 exactly like
    the upstream kernel, accumulates in fp32 in TMEM, rounds the result once)."""
    w16 = dequantize(qweight, qzeros, scales, group, np.float16).astype(np.float64)
    y = np.asarray(x, dtype=np.float16).astype(np.float64) @ w16
    return y.astype(np.float16)


def gemm_reference_emulated(x: np.ndarray, qweight, qzeros, scales, group: int,
                            split_k_iters: int = 8) -> np.ndarray:
    """Emulation of ``awq_ext.gemm_forward_cuda(..., split_k_iters=8)`` rounding (see module
    docstring).  Used only to measure how far the *reference itself* sits from gemm_exact so that
    the parity tolerance can be stated relative to it."""
    x32 = np.asarray(x, dtype=np.float16).astype(np.float32)
    w16 = dequantize(qweight, qzeros, scales, group, np.float16).astype(np.float32)
    k = w16.shape[0]
    # upstream: k-tiles of 32 are dealt round-robin to the split_k_iters slices
    tiles = np.arange(k // 32)
    parts = []
    for s in range(split_k_iters):
        rows = np.concatenate([np.arange(t * 32, t * 32 + 32) for t in tiles[tiles % split_k_iters == s]]) \
            if np.any(tiles % split_k_iters == s) else np.zeros(0, dtype=np.int64)
        if len(rows) == 0:
            parts.append(np.zeros((x32.shape[0], w16.shape[1]), dtype=np.float16))
            continue
        parts.append((x32[:, rows] @ w16[rows]).astype(np.float16))
    return np.sum(np.stack(parts).astype(np.float32), axis=0).astype(np.float16)


def random_awq(k: int, n: int, group: int, rng: np.random.Generator, scale: float = 0.003):
    """Synthetic AWQ triplet with the statistics SURVEY.md 8(d) prescribes: uniform nibbles,
    scales = scale * (0.5 + U[0,1)) in fp16."""
    qw = rng.integers(0, 16, size=(k, n), dtype=np.int64)
    qz = rng.integers(0, 16, size=(k // group, n), dtype=np.int64)
    sc = (scale * (0.5 + rng.random((k // group, n)))).astype(np.float16)
    return pack_cols_awq(qw), pack_cols_awq(qz), sc


def dequantize_torch(qweight, qzeros, scales, group: int, exact: bool = False):
    """Device-agnostic torch twin of ``dequantize`` (same nibble order): fp32 [K, N] holding either the
    fp16-rounded weights fp16((q - z) * s) -- what both upstream kernels feed their MMAs -- or, with
    ``exact=True``, the unrounded product (exact in fp32: a 5-bit integer times an 11-bit mantissa).
    Used where the numpy version is too slow (70B-shape parity tests run the oracle on the GPU box)."""
    import torch
    shifts = torch.tensor([4 * int(p) for p in AWQ_NIBBLE_OF_COL], dtype=torch.int32, device=qweight.device)
    q = ((qweight.unsqueeze(-1) >> shifts) & 0xF).reshape(qweight.shape[0], -1)            # [K, N]
    z = ((qzeros.unsqueeze(-1) >> shifts) & 0xF).reshape(qzeros.shape[0], -1)              # [K/g, N]
    k = q.shape[0]
    d = (q - z.repeat_interleave(group, dim=0)[:k]).to(torch.float32)
    w = d * scales.to(torch.float32).repeat_interleave(group, dim=0)[:k]
    return w if exact else w.to(torch.float16).to(torch.float32)

"""CPU baseline of the verify path: the oracle's arithmetic timed on host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Used by ``bench.py``'s
``cpu_baseline`` leg and by ``bench.py --impl reference``; never by the product path.

What is timed: one decoder layer of the Llama-3.1-70B-AWQ verify forward over the N tree tokens, as
a CPU implementation of the reference's path has to do it (reference call sites:
umbrella/models/llama.py:262-303, umbrella/quantization/awq_utils.py:63-86).  The seven INT4 linears --
99.9 % of the arithmetic -- run in the C restatement oracle/awq_cpu.c (weights stay INT4 in memory, one pass
over them with unit stride, a 16-entry fp16((q - z) * s) table per column and group, pthreads over column
blocks; no dequantised matrix is materialised); RMSNorm, RoPE, masked GQA attention over the prefix and SiLU
are torch CPU ops.  The reference itself has no CPU path (it is GPU compute with optional host-RAM weight
paging); this port is the closest CPU statement of the same arithmetic.  A full verify is 80
such layers + the fp16 lm_head; the sample is ONE layer (+ a 1/80 share of the lm_head), scaled by 80 and
stated as such.
"""
from __future__ import annotations

import math
import time

import torch

import ctypes as C
import os
import subprocess

from . import llama as ollama

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libawq_cpu.so")
_lib = None
_THREADS = [1]


def build_c_oracle(force: bool = False) -> str:
    """make -C oracle (gcc, x86-64-v3, pthreads) -> oracle/_build/libawq_cpu.so."""
    src = os.path.join(HERE, "awq_cpu.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB


def c_oracle():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c_oracle())
    return _lib


def awq_gemm_c(x16: torch.Tensor, qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, group: int,
               threads: int | None = None) -> torch.Tensor:
    """fp16(x @ fp16((q - z) * s)) through oracle/awq_cpu.c.  x16 [M, K] fp16 (any M)."""
    x16 = x16.contiguous()
    m, k = x16.shape
    n = scales.shape[1]
    y = torch.empty((m, n), dtype=torch.float32)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = c_oracle().oracle_awq_gemm(p(x16), p(qweight.contiguous()), p(qzeros.contiguous()), p(scales.contiguous()), m, k, n,
                                    group, p(y), threads or _THREADS[0])
    assert rc == 0, rc
    return y.to(torch.float16)


def _dequant_matmul(x16, qweight, qzeros, scales, group):
    return awq_gemm_c(x16, qweight, qzeros, scales, group)


class CpuLayerBaseline:
    def __init__(self, shape: ollama.LlamaShape = ollama.LLAMA31_70B, tree_size: int = 13, prefix: int = 512,
                 seed: int = 1234, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.threads = threads or os.cpu_count() or torch.get_num_threads()
        _THREADS[0] = self.threads
        c_oracle()
        self.shape, self.N, self.prefix = shape, tree_size, prefix
        self.w = ollama.synth_layer(shape, 0, awq=True, seed=seed)
        g = torch.Generator().manual_seed(seed)
        self.hidden = torch.randn(1, tree_size, shape.hidden_size, generator=g).to(torch.float16)
        hkv, d = shape.num_key_value_heads, shape.head_dim
        self.k = torch.randn(prefix + tree_size, hkv, d, generator=g).to(torch.float16)
        self.v = torch.randn(prefix + tree_size, hkv, d, generator=g).to(torch.float16)
        inv, sc = ollama.llama3_inv_freq(shape)
        self.cos, self.sin = ollama.rope_tables(inv, sc, prefix + tree_size + 8)
        self.pos = (prefix + torch.arange(tree_size))[None]
        self.mask = torch.ones(tree_size, prefix + tree_size, dtype=torch.bool)
        self.mask[:, prefix:] = torch.tril(torch.ones(tree_size, tree_size, dtype=torch.bool))
        self.lm_head_rows = shape.vocab_size // shape.num_hidden_layers        # 1/80 share of the lm_head
        self.lm_head = (0.02 * torch.randn(self.lm_head_rows, shape.hidden_size, generator=g)).to(torch.float16)

    @torch.inference_mode()
    def layer_pass(self) -> float:
        """One layer of the verify forward on CPU; returns seconds."""
        sh, w, N = self.shape, self.w, self.N
        gsz = sh.group_size
        t0 = time.perf_counter()
        x = ollama.rmsnorm(self.hidden, w["input_layernorm"], sh.rms_norm_eps)[0]
        q = _dequant_matmul(x, *w["wq"], gsz).view(1, N, sh.num_attention_heads, sh.head_dim)
        k = _dequant_matmul(x, *w["wk"], gsz).view(1, N, sh.num_key_value_heads, sh.head_dim)
        v = _dequant_matmul(x, *w["wv"], gsz).view(1, N, sh.num_key_value_heads, sh.head_dim)
        q, k = ollama.apply_rope(q, k, self.cos, self.sin, self.pos)
        self.k[self.prefix:] = k[0]
        self.v[self.prefix:] = v[0]
        g = sh.num_attention_heads // sh.num_key_value_heads
        qq = q[0].float().reshape(N, sh.num_key_value_heads, g, -1)
        s = torch.einsum("qhgd,khd->hgqk", qq, self.k.float()) / math.sqrt(sh.head_dim)
        s = s.masked_fill(~self.mask[None, None], float("-inf"))
        o = torch.einsum("hgqk,khd->qhgd", torch.softmax(s, dim=-1), self.v.float()).reshape(N, -1).to(torch.float16)
        h = self.hidden[0] + _dequant_matmul(o, *w["wo"], gsz)
        x = ollama.rmsnorm(h, w["post_attention_layernorm"], sh.rms_norm_eps)
        up = _dequant_matmul(x, *w["up_proj"], gsz)
        gate = torch.nn.functional.silu(_dequant_matmul(x, *w["gate_proj"], gsz).float()).to(torch.float16)
        h = h + _dequant_matmul(gate * up, *w["down_proj"], gsz)
        _ = (h.float() @ self.lm_head.float().t()).to(torch.float16)
        return time.perf_counter() - t0

    def tokens_per_second(self, layer_seconds: float, accept_len: float) -> float:
        """Extrapolated accepted tokens/sec of a CPU verify loop: 80 layers per verify (the draft's cost
        is left out, which flatters the CPU arm)."""
        return accept_len / (layer_seconds * self.shape.num_hidden_layers)

"""GPU parity: glue kernels (RMSNorm, RoPE + KV append, SiLU*up, add, embedding, arg-max) vs the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import llama as ollama

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def test_rmsnorm_bit_exact_vs_oracle():
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(0)
    for rows, hidden in [(1, 2048), (13, 8192), (31, 4096), (128, 8192)]:
        x = torch.randn(rows, hidden, generator=g).to(torch.float16)
        w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.float16)
        ref = ollama.rmsnorm(x, w, 1e-5)
        out = ops.rmsnorm(x.to(DEV), w.to(DEV), 1e-5).cpu()
        # fp32 reduction order differs from torch's: allow 1 fp16 ulp on a handful of elements
        diff = (out.float() - ref.float()).abs()
        assert diff.max() <= 2 * torch.finfo(torch.float16).eps * ref.float().abs().max()
        assert (out != ref).float().mean() < 0.01


def test_rope_golden_and_append():
    """rope.npz was produced by the reference's apply_rotary_pos_emb (oracle/make_golden.py)."""
    from umbrella_b200 import ops
    d = np.load(os.path.join(GOLDEN, "rope.npz"))
    cos, sin = torch.from_numpy(d["rope_cos"]).to(DEV), torch.from_numpy(d["rope_sin"]).to(DEV)
    q, k = torch.from_numpy(d["rope_q"])[0], torch.from_numpy(d["rope_k"])[0]          # [13, Hq, D], [13, Hkv, D]
    pos = torch.from_numpy(d["rope_pos"])[0].to(torch.int64)
    qo, Hq, D = q.shape
    Hkv = k.shape[1]
    v = torch.randn(qo, Hkv, D).to(torch.float16)
    qkv = torch.cat([q.reshape(qo, -1), k.reshape(qo, -1), v.reshape(qo, -1)], dim=1).contiguous().to(DEV)
    S, off = 64, 17
    for layout in ("NHD", "HND"):
        kc = torch.zeros((S, Hkv, D) if layout == "NHD" else (Hkv, S, D), dtype=torch.float16, device=DEV)
        vc = torch.zeros_like(kc)
        q_out = torch.empty(qo, Hq, D, dtype=torch.float16, device=DEV)
        ops.rope_append(qkv, q_out, kc, vc, cos, sin, pos.to(DEV), Hq, Hkv, D, off, layout)
        torch.cuda.synchronize()
        kk = kc[off:off + qo] if layout == "NHD" else kc[:, off:off + qo].transpose(0, 1)
        vv = vc[off:off + qo] if layout == "NHD" else vc[:, off:off + qo].transpose(0, 1)
        assert torch.equal(q_out.cpu(), torch.from_numpy(d["rope_q_out"])[0])
        assert torch.equal(kk.cpu(), torch.from_numpy(d["rope_k_out"])[0])
        assert torch.equal(vv.cpu(), v)
        assert float(kc.abs().sum() - kk.abs().sum()) == 0.0     # nothing written outside the appended rows


def test_silu_add_embedding():
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(1)
    gate = (2 * torch.randn(13, 28672, generator=g)).to(torch.float16)
    up = torch.randn(13, 28672, generator=g).to(torch.float16)
    ref = torch.nn.functional.silu(gate.float()).to(torch.float16) * up
    out = ops.silu_mul(gate.to(DEV), up.to(DEV)).cpu()
    assert (out.float() - ref.float()).abs().max() <= 2e-3 * ref.float().abs().max()
    assert (out != ref).float().mean() < 0.01           # expf vs torch's exp: <= 1 ulp on rare elements
    a, b = torch.randn(13, 8192, generator=g).to(torch.float16), torch.randn(13, 8192, generator=g).to(torch.float16)
    assert torch.equal(ops.add(a.to(DEV), b.to(DEV)).cpu(), a + b)
    table = torch.randn(1000, 512, generator=g).to(torch.float16)
    ids = torch.randint(0, 1000, (31,), generator=g)
    assert torch.equal(ops.embedding(ids.to(DEV), table.to(DEV)).cpu(), table[ids])


def test_argmax_first_occurrence():
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(13, 128256, generator=g)
    logits[3, 77] = logits[3, 90000] = 9.0          # tie -> first index
    logits[5] = -1e30
    logits[5, 128255] = -5.0
    out = ops.argmax(logits.to(DEV)).cpu()
    ref = logits.argmax(-1)
    assert torch.equal(out, ref) and out[3] == 77
    small = torch.randn(4, 100, generator=g)
    assert torch.equal(ops.argmax(small.to(DEV)).cpu(), small.argmax(-1))

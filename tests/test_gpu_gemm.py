"""GPU parity: AWQ INT4 repack / dequantise / tcgen05 GEMM and the fp16 GEMM vs the numpy oracle
(oracle/awq.py).  Tolerances: the kernel rounds (q - z) * s to fp16 exactly like the reference
kernels (bit-exact, checked through the unpack), accumulates in fp32 in tensor memory and rounds the
result once to fp16, so it must sit within a few fp16 ulps of ``gemm_fp16_weights`` -- far inside
north_star's 1e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import awq as oawq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel_err(out, ref):
    ref = ref.astype(np.float64)
    return np.abs(out.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-6)


def _awq(K, N, seed):
    rng = np.random.default_rng(seed)
    qw, qz, sc = oawq.random_awq(K, N, 128, rng)
    return qw, qz, sc, rng


@pytest.mark.parametrize("mode", [0, 1])
def test_probe_umma(mode):
    """One 128x16x64 UMMA: A from shared memory (0) / tensor memory (1) -- pins the descriptor and
    TMEM operand layouts the GEMM kernels rely on."""
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(mode)
    a = torch.randn(128, 64, generator=g).to(torch.float16)
    b = torch.randn(16, 64, generator=g).to(torch.float16)
    d = ops.probe_umma(a.to(DEV), b.to(DEV), mode).cpu()
    ref = a.float() @ b.float().t()
    assert torch.allclose(d, ref, atol=1e-3, rtol=1e-4), (d - ref).abs().max()


def test_probe_umma_mn_major_operand():
    """tcgen05.mma with an MN-major B operand (the way V rows lie in a KV cache): LBO = stride between 64-element atoms
    along N, SBO = 1024 B between 8-row groups along K.  The swapped reading (mode 3) must be wrong."""
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(7)
    a = torch.randn(128, 64, generator=g).to(torch.float16)
    bt = torch.randn(64, 128, generator=g).to(torch.float16)          # [k, n], n contiguous
    ref = a.float() @ bt.float()
    d = ops.probe_umma(a.to(DEV), bt.to(DEV), 2).cpu()
    assert d.shape == (128, 128) and torch.allclose(d, ref, atol=1e-3, rtol=1e-4), (d - ref).abs().max()
    assert (ops.probe_umma(a.to(DEV), bt.to(DEV), 3).cpu() - ref).abs().max() > 1.0


def test_repack_unpack_dequantize_bit_exact():
    from umbrella_b200 import ops
    golden = np.load(os.path.join(GOLDEN, "awq_pack.npz"))
    assert np.array_equal(oawq.pack_cols_awq(golden["q2"]), golden["packed2"])       # format pinned (vLLM awq_pack)
    for K, N, seed in [(128, 256, 0), (512, 1024, 1), (1024, 384, 2)]:
        qw, qz, sc, _ = _awq(K, N, seed)
        w16 = oawq.dequantize(qw, qz, sc, 128, np.float16)                             # [K, N]
        t = [torch.from_numpy(a).to(DEV) for a in (qw, qz, sc)]
        assert np.array_equal(ops.awq_dequantize(*t).cpu().numpy().view(np.uint16), w16.view(np.uint16))
        n_total = (N + 127) // 128 * 128
        packed = ops.awq_repack(*t, N_total=n_total)
        un = ops.awq_unpack(packed, K, n_total).cpu().numpy()                          # [N_total, K]
        assert np.array_equal(un[:N].view(np.uint16), np.ascontiguousarray(w16.T).view(np.uint16))
        assert not un[N:].any()


def test_repack_fused_rows():
    """q|k|v concatenation and the 64|64 gate/up interleave share one stream."""
    from umbrella_b200 import ops
    K = 256
    a, b = _awq(K, 256, 3)[:3], _awq(K, 128, 4)[:3]
    ta = [torch.from_numpy(x).to(DEV) for x in a]
    tb = [torch.from_numpy(x).to(DEV) for x in b]
    packed = ops.awq_repack(*ta, N_total=384)
    ops.awq_repack(*tb, packed=packed, N_total=384, row_offset=256)
    un = ops.awq_unpack(packed, K, 384).cpu().numpy()
    wa, wb = oawq.dequantize(*a, 128, np.float16).T, oawq.dequantize(*b, 128, np.float16).T
    assert np.array_equal(un[:256], wa) and np.array_equal(un[256:], wb)
    g, u = _awq(K, 256, 5)[:3], _awq(K, 256, 6)[:3]
    packed = ops.awq_repack(*[torch.from_numpy(x).to(DEV) for x in g], N_total=512, row_block=64, row_stride=128)
    ops.awq_repack(*[torch.from_numpy(x).to(DEV) for x in u], packed=packed, N_total=512, row_block=64, row_stride=128,
                   row_offset=64)
    un = ops.awq_unpack(packed, K, 512).cpu().numpy().reshape(4, 2, 64, K)
    wg, wu = oawq.dequantize(*g, 128, np.float16).T.reshape(4, 64, K), oawq.dequantize(*u, 128, np.float16).T.reshape(4, 64, K)
    assert np.array_equal(un[:, 0], wg) and np.array_equal(un[:, 1], wu)


@pytest.mark.parametrize("M,K,N", [(13, 512, 256), (1, 1024, 1024), (16, 8192, 1024), (31, 4096, 4096),
                                    (64, 2048, 640), (100, 1024, 2048), (128, 512, 128), (13, 28672, 512)])
def test_awq_gemm_vs_oracle(M, K, N):
    from umbrella_b200 import ops
    qw, qz, sc, rng = _awq(K, N, M + K)
    x = rng.standard_normal((M, K)).astype(np.float16)
    ref = oawq.gemm_fp16_weights(x, qw, qz, sc, 128)
    packed = ops.awq_repack(*[torch.from_numpy(a).to(DEV) for a in (qw, qz, sc)])
    xd = torch.from_numpy(x).to(DEV)
    out = ops.awq_gemm(xd, packed, K, N).cpu().numpy()
    assert _rel_err(out, ref) < 1e-3
    simt = ops.awq_gemm_simt(xd, packed, K, N).cpu().numpy()
    assert _rel_err(simt, ref) < 1e-3
    # determinism: stream-K partials are reduced in a fixed order
    out2 = ops.awq_gemm(xd, packed, K, N).cpu().numpy()
    assert np.array_equal(out.view(np.uint16), out2.view(np.uint16))


def test_awq_gemm_epilogues():
    from umbrella_b200 import ops
    M, K, N = 13, 1024, 512
    qw, qz, sc, rng = _awq(K, N, 9)
    x = rng.standard_normal((M, K)).astype(np.float16)
    y = oawq.gemm_fp16_weights(x, qw, qz, sc, 128)
    t = [torch.from_numpy(a).to(DEV) for a in (qw, qz, sc)]
    xd = torch.from_numpy(x).to(DEV)
    packed = ops.awq_repack(*t)
    res = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float16)).to(DEV)
    out = ops.awq_gemm(xd, packed, K, N, mode=ops.UMB_EPI_RESIDUAL, residual=res)
    ref = torch.from_numpy(y) + res.cpu()
    assert _rel_err(out.cpu().numpy(), ref.numpy()) < 2e-3
    inplace = res.clone()
    ops.awq_gemm(xd, packed, K, N, mode=ops.UMB_EPI_RESIDUAL, residual=inplace, out=inplace)
    assert torch.equal(inplace, out)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float16)).to(DEV)
    outb = ops.awq_gemm(xd, packed, K, N, bias=bias)
    assert _rel_err(outb.cpu().numpy(), (torch.from_numpy(y) + bias.cpu()).numpy()) < 2e-3
    # SiLU * up: tile = 64 gate | 64 up rows
    gq, uq = _awq(K, 256, 10)[:3], _awq(K, 256, 11)[:3]
    packed2 = ops.awq_repack(*[torch.from_numpy(a).to(DEV) for a in gq], N_total=512, row_block=64, row_stride=128)
    ops.awq_repack(*[torch.from_numpy(a).to(DEV) for a in uq], packed=packed2, N_total=512, row_block=64,
                   row_stride=128, row_offset=64)
    g16 = torch.from_numpy(oawq.gemm_fp16_weights(x, *gq, 128))
    u16 = torch.from_numpy(oawq.gemm_fp16_weights(x, *uq, 128))
    ref = torch.nn.functional.silu(g16.float()).to(torch.float16) * u16
    out = ops.awq_gemm(xd, packed2, K, 512, mode=ops.UMB_EPI_SILU_MUL)
    assert out.shape == (M, 256)
    assert _rel_err(out.cpu().numpy(), ref.numpy()) < 3e-3


@pytest.mark.parametrize("M,K,N", [(13, 512, 1024), (1, 2048, 4096), (31, 8192, 1280), (128, 1024, 256), (5, 2048, 128256)])
def test_gemm_f16_and_logits(M, K, N):
    from umbrella_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + 1)
    x = torch.randn(M, K, generator=g).to(torch.float16)
    w = (0.02 * torch.randn(N, K, generator=g)).to(torch.float16)
    ref = (x.double() @ w.double().t())
    xd, wd = x.to(DEV), w.to(DEV)
    out = ops.gemm_f16(xd, wd).cpu()
    assert _rel_err(out.numpy(), ref.numpy()) < 1e-3
    keys = torch.zeros(M, dtype=torch.int64, device=DEV)
    logits = ops.gemm_f16(xd, wd, mode=ops.UMB_EPI_LOGITS, argmax=keys)
    assert logits.dtype == torch.float32
    assert torch.equal(logits.cpu(), out.float())                   # float(fp16(acc)), llama.py:321
    ids = ops.argmax_keys_to_ids(keys).cpu()
    assert torch.equal(ids, logits.cpu().argmax(-1))
    assert torch.equal(ops.argmax(logits).cpu(), ids)
    assert int(keys.abs().sum()) == 0                               # key buffer handed back clean


def test_awq_kernels_vs_upstream_golden():
    """umb_awq_dequantize_f16 is bit-identical to the upstream dequantize_weights kernel, and the tcgen05 GEMM
    is at least as close to the exact product as the upstream gemm_forward kernel (tests/golden/awq_upstream.npz,
    produced by oracle/make_golden_gpu.py from vLLM's build of the autoawq-kernels sources)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "awq_upstream.npz")
    if not os.path.exists(path):
        pytest.skip("awq_upstream.npz not generated yet")
    from umbrella_b200 import ops
    z = np.load(path)
    ci = 0
    while "c%d_meta" % ci in z:
        pre = "c%d_" % ci
        m, k, n, g = map(int, z[pre + "meta"])
        qw, qz, sc, x = (torch.from_numpy(z[pre + s]).to(DEV) for s in ("qweight", "qzeros", "scales", "x"))
        deq = ops.awq_dequantize(qw, qz, sc)
        assert torch.equal(deq.view(torch.int16).cpu(), torch.from_numpy(z[pre + "dequant"]).view(torch.int16)), ci
        packed = ops.awq_repack(qw, qz, sc)
        y = ops.awq_gemm(x, packed, k, n).float().cpu().numpy().astype(np.float64)
        exact = oawq.gemm_exact(z[pre + "x"], z[pre + "qweight"], z[pre + "qzeros"], z[pre + "scales"], g)
        scale = np.abs(exact).max()
        err_up = np.abs(z[pre + "y"].astype(np.float64) - exact).max() / scale
        err = np.abs(y - exact).max() / scale
        assert err <= max(err_up, 1e-3), (ci, err, err_up)          # tolerance: no worse than the reference kernel
        ci += 1
    assert ci >= 4


@pytest.mark.parametrize("awq,M,Hq,Hkv,D,K", [(True, 13, 8, 2, 128, 1024), (True, 3, 4, 1, 64, 512), (False, 31, 4, 2, 64, 256),
                                             (True, 97, 4, 4, 128, 512), (False, 1, 8, 8, 128, 512)])
def test_fused_rope_qkv_epilogue_is_bit_identical(awq, M, Hq, Hkv, D, K):
    """UMB_EPI_ROPE_QKV (q|k|v GEMM + RoPE + KV append in one kernel) writes exactly the bits of the two-kernel
    path (GEMM store, then umb_rope_append_f16): same fp16 rounding points, same three rounded rotary ops."""
    from umbrella_b200 import ops
    rng = np.random.default_rng(M + D)
    N, S, off = (Hq + 2 * Hkv) * D, 256, 77
    x = torch.from_numpy((rng.standard_normal((M, K)) * 0.7).astype(np.float16)).to(DEV)
    if awq:
        qw, qz, sc = oawq.random_awq(K, N, 128, rng, scale=0.02)
        packed = ops.awq_repack(*(torch.from_numpy(t).to(DEV) for t in (qw, qz, sc)))
        gemm = lambda **kw: ops.awq_gemm(x, packed, K, N, **kw)
    else:
        w = torch.from_numpy((rng.standard_normal((N, K)) * 0.05).astype(np.float16)).to(DEV)
        gemm = lambda **kw: ops.gemm_f16(x, w, **kw)
    pos = torch.from_numpy(rng.integers(0, S, M)).to(DEV)
    ang = torch.rand(S, D, device=DEV) * 6.28
    cos, sin = ang.cos().to(torch.float16), ang.sin().to(torch.float16)
    outs = []
    for fused in (False, True):
        q_out = torch.zeros(M, Hq, D, dtype=torch.float16, device=DEV)
        kc = torch.zeros(S, Hkv, D, dtype=torch.float16, device=DEV)
        vc = torch.zeros_like(kc)
        if fused:
            rp = ops.make_rope_epilogue(q_out, kc, vc, cos, sin, pos, Hq, Hkv, D, off)
            gemm(mode=ops.UMB_EPI_ROPE_QKV, rope=rp, out=torch.zeros(M, N, dtype=torch.float16, device=DEV))
        else:
            qkv = gemm()
            ops.rope_append(qkv, q_out, kc, vc, cos, sin, pos, Hq, Hkv, D, off)
        outs.append((q_out, kc, vc))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert outs[1][1][off:off + M].abs().sum() > 0 and outs[1][1][:off].abs().sum() == 0


@pytest.mark.parametrize("M", [65, 97, 128])
def test_rows_above_64_are_stable_and_correct(M):
    """Regression for the rare launch failure of the 128-column GEMM instantiation (scripts/repro_m128.py; root cause:
    a parity wait two phases ahead of its barrier in the 4-deep ring, profiles/r2_m128_root_cause.txt -- 24 000 launches
    clean after the fix, one in ~150 failing before): 200 back-to-back single launches of 65..128 rows at a 70B q|k|v
    shape must all succeed, leave the workspace clean and agree with the first result bit for bit (deterministic
    reduction)."""
    from umbrella_b200 import ops
    K, N = 8192, 10240
    rng = np.random.default_rng(M)
    qw, qz, sc = oawq.random_awq(K, N, 128, rng)
    packed = ops.awq_repack(*(torch.from_numpy(t).to(DEV) for t in (qw, qz, sc)))
    x = torch.from_numpy((rng.standard_normal((M, K)) * 0.5).astype(np.float16)).to(DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    first = ops.awq_gemm(x, packed, K, N, workspace=ws).clone()
    for _ in range(200):
        y = ops.awq_gemm(x, packed, K, N, workspace=ws)
    torch.cuda.synchronize()
    assert torch.equal(y, first)
    assert int(ws.view(torch.int32).abs().max()) == 0
    w16 = ops.awq_unpack(packed, K, N).float()                       # [N, K] fp16 weights as the kernel sees them
    ref = x.float() @ w16.t()
    assert float((y.float() - ref).abs().max() / ref.abs().max()) < 2e-3

"""CPU: the sampling oracle (oracle/sampling.py) -- internal consistency of the restated flashinfer algorithms, and
the product's torch fallback for topk > 1024 against it."""
import numpy as np
import torch

from oracle import sampling as S


def test_nucleus_is_top_p_renorm_kept_set_with_ties():
    p = np.array([0.3, 0.3, 0.2, 0.1, 0.1])
    assert S.nucleus(p, 0.5).tolist() == [True, True, False, False, False]      # both 0.3s: mass above them is 0
    assert S.nucleus(p, 0.61).tolist() == [True, True, True, False, False]
    assert S.nucleus(p, 0.85).tolist() == [True] * 5                            # the tied 0.1s enter together
    assert S.nucleus(np.array([1.0, 0.0, 0.0]), 0.9).tolist() == [True, False, False]


def test_top_k_mask_keeps_ties_at_the_kth_value():
    x = np.array([[5.0, 3.0, 3.0, 3.0, 1.0]])
    assert np.isfinite(S.top_k_mask(x, 2)).tolist() == [[True, True, True, True, False]]
    assert np.isfinite(S.top_k_mask(x, 99)).all()


def test_both_samplers_draw_from_the_renormalised_nucleus():
    rng = np.random.default_rng(0)
    row = rng.standard_normal(1000) * 3
    dist = S.exact_distribution(row, 0.6, 32, 0.9)
    M = 20000
    us = rng.random((3, M))
    rows = np.repeat(row[None], M, 0)
    for out in (S.sample_static(rows, us, 0.6, 32, 0.9)[0], S.sample_dynamic(rows, us[0], 0.6, 32, 0.9)[0]):
        cnt = np.bincount(out, minlength=1000) / M
        assert 0.5 * np.abs(cnt - dist).sum() < 0.03


def test_repetition_penalty_is_applied_once_per_token_id():
    x = np.array([[2.0, -2.0, 1.0]])
    out = S.apply_repetition_penalty(np.array([0, 0, 1, 1, 1]), x, 2.0)
    assert out.tolist() == [[1.0, -4.0, 1.0]]
    assert S.argmax_penalised(x, np.array([0]), 4.0).tolist() == [2]


def test_torch_fallback_sampler_matches_oracle():
    from umbrella_b200.speculation.speculation_utils import sample_topk_topp_torch
    rng = np.random.default_rng(1)
    logits = (rng.standard_normal((64, 3000)) * 3).astype(np.float16).astype(np.float32)     # fp16 grid: ties
    u = rng.random((3, 64)).astype(np.float32)
    pen = rng.integers(0, 3000, 200)
    for static in (True, False):
        o, m = (S.sample_static if static else S.sample_dynamic)(logits, u if static else u[0], 0.6, 2000, 0.9, pen, 1.1)
        t = sample_topk_topp_torch(torch.from_numpy(logits), 0.6, 2000, 0.9, torch.from_numpy(u if static else u[0]),
                                   static, torch.from_numpy(pen), 200, 1.1).numpy()
        assert ((o == t) | (m < 1e-5)).all()

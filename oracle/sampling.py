"""Oracle: the target-side sampling of verify() (repetition penalty, top-k, softmax(T), top-p, one sample per
tree node) in numpy float64.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (paths relative to /root/reference):

  * ``apply_repetition_penalty``  umbrella/speculation/speculation_utils.py:340-345
  * ``apply_topk``                umbrella/speculation/speculation_utils.py:347-352
  * static engine, T >= 0.05      umbrella/speculation/static_speculation_engine.py:305-310
        sampled, _ = flashinfer.sampling.top_k_top_p_sampling_from_logits(
            target_logits / T, self.uniform_samples, self.topk, self.topp)
    with ``self.uniform_samples = torch.rand(3, N)`` drawn ONCE at initialize() (:131): the sample is a
    deterministic function of the logits.
  * dynamic engine, T >= 0.05     umbrella/speculation/dynamic_speculation_engine.py:276-281
        logits = apply_topk(logits, topk); proba = softmax(logits / T)
        proba = flashinfer.sampling.top_p_renorm_prob(proba, topp); sampled = torch.multinomial(proba, 1)

The flashinfer functions live in a third-party dependency the reference does not pin (its README installs a
flashinfer 0.1.x wheel; the 4-argument ``(logits, uniform_samples, top_k, top_p)`` form with a
``[max_rounds, batch]`` uniform tensor exists in flashinfer 0.1.x - 0.2.0 only).  Their published algorithm
(include/flashinfer/sampling.cuh of those releases; the top-p renormalisation kernel is unchanged in the 0.6.11
copy in this image, sampling.cuh:1661-1830) is restated here:

  ``top_k_top_p_sampling_from_logits`` (default filter_apply_order "top_k_first")
      = top_k_mask_logits -> softmax -> top_p_sampling_from_probs
  ``top_k_mask_logits``: logits below the k-th largest VALUE become -inf (ties at the k-th value all stay).
  ``top_p_sampling_from_probs(probs, uniform[R, B], p)`` -- rejection sampling without sorting:
      q = 1, pivot = 0
      for r in range(R):
          u = uniform[r, b] * q
          sampled = first index j (ascending token id) whose running sum of probs[probs > pivot] exceeds u
          pivot = max(pivot, probs[sampled]);  q = sum(probs[probs > pivot])
          if q < p: break                       # the sampled token lies inside the top-p nucleus
      return sampled                            # after R failed rounds: the last draw (success = False)
  ``top_p_renorm_prob(probs, p)``: keep the smallest set {probs >= v} whose mass reaches p -- token j stays iff
      sum(probs[probs > probs[j]]) < p -- and renormalise.

``torch.multinomial`` draws from torch's Philox stream, which nothing outside torch can reproduce; the dynamic
path is therefore restated with an explicit uniform per node (inverse CDF in ascending token id over the
renormalised nucleus) -- same distribution, deterministic given the uniforms.  That substitution is the one
deliberate difference and is what the product implements as well.

No golden vector exists for either flashinfer function in the reference -> "parity unpinned" for the sampling
arithmetic; the distribution is pinned on the GPU box against flashinfer 0.6.11's own
``top_k_top_p_sampling_from_logits`` / ``top_p_renorm_probs`` (tests/test_gpu_sampling.py).
"""
from __future__ import annotations

import numpy as np


def apply_repetition_penalty(input_ids: np.ndarray, logits: np.ndarray, penalty: float) -> np.ndarray:
    """speculation_utils.py:340-345: gather the logits of the ids seen so far, scale them away from the
    arg-max (x * p if x < 0 else x / p), scatter back.  ``input_ids`` [n] is shared by all rows."""
    out = np.array(logits, dtype=np.float64, copy=True)
    ids = np.unique(np.asarray(input_ids, dtype=np.int64))
    g = out[:, ids]
    out[:, ids] = np.where(g < 0, g * penalty, g / penalty)
    return out


def top_k_mask(logits: np.ndarray, k: int) -> np.ndarray:
    """apply_topk (speculation_utils.py:347-352) / flashinfer top_k_mask_logits: values below the k-th largest
    value of the row become -inf; ties at the k-th value stay."""
    x = np.asarray(logits, dtype=np.float64)
    k = min(int(k), x.shape[-1])
    kth = np.partition(x, x.shape[-1] - k, axis=-1)[..., x.shape[-1] - k, None]
    return np.where(x < kth, -np.inf, x)


def softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def nucleus(probs: np.ndarray, top_p: float) -> np.ndarray:
    """bool mask of top_p_renorm_prob's kept set: sum(probs[probs > probs[j]]) < top_p."""
    p = np.asarray(probs, dtype=np.float64)
    order = np.argsort(-p, kind="stable")
    sp = p[order]
    # mass strictly greater than each value: cumulative sum before the first occurrence of that value
    csum = np.concatenate([[0.0], np.cumsum(sp)[:-1]])
    first = np.r_[True, sp[1:] != sp[:-1]]
    gt_mass = np.maximum.accumulate(np.where(first, csum, 0.0))
    keep = np.zeros_like(p, dtype=bool)
    keep[order] = gt_mass < top_p
    return keep & (p > 0)


def top_p_sampling_from_probs(probs: np.ndarray, uniform: np.ndarray, top_p: float) -> tuple[int, bool, float]:
    """One row of flashinfer 0.1.x top_p_sampling_from_probs.  uniform: [R].  Returns (sampled, success, margin)
    where margin is the smallest |running sum - u| met at a decision (tests skip rows whose margin is inside the
    fp32 noise of the device kernel)."""
    p = np.asarray(probs, dtype=np.float64)
    q, pivot, sampled, margin = 1.0, 0.0, p.size - 1, np.inf
    for r in range(len(uniform)):
        u = float(uniform[r]) * q
        masked = np.where(p > pivot, p, 0.0)
        cdf = np.cumsum(masked)
        hit = np.nonzero(cdf > u)[0]
        sampled = int(hit[0]) if hit.size else p.size - 1
        live = np.nonzero(masked > 0)[0]
        margin = min(margin, float(np.abs(cdf[live] - u).min())) if live.size else margin
        pivot = max(pivot, p[sampled])
        q = float(p[p > pivot].sum())
        margin = min(margin, abs(q - top_p))
        if q < top_p:
            return sampled, True, margin
    return sampled, False, margin


def sample_static(logits: np.ndarray, uniform_samples: np.ndarray, temperature: float, top_k: int, top_p: float,
                  penalty_ids=None, penalty: float = 1.0):
    """static_speculation_engine.py:298-310 for T >= 0.05 -> (sampled int64 [N], margin float [N])."""
    x = np.asarray(logits, dtype=np.float64)
    if penalty > 1.01 and penalty_ids is not None:
        x = apply_repetition_penalty(penalty_ids, x, penalty)
    probs = softmax(top_k_mask(x / temperature, top_k))     # float64 throughout; the device works in fp32 -> `margin`
    out, margins = [], []
    for i in range(x.shape[0]):
        s, _, m = top_p_sampling_from_probs(probs[i], np.asarray(uniform_samples)[:, i], top_p)
        out.append(s)
        margins.append(m)
    return np.asarray(out, dtype=np.int64), np.asarray(margins)


def sample_dynamic(logits: np.ndarray, uniform: np.ndarray, temperature: float, top_k: int, top_p: float,
                   penalty_ids=None, penalty: float = 1.0):
    """dynamic_speculation_engine.py:265-281 for T >= 0.05 with an explicit uniform per node in place of
    torch.multinomial's private random stream -> (sampled int64 [N], margin float [N])."""
    x = np.asarray(logits, dtype=np.float64)
    if penalty > 1.01 and penalty_ids is not None:
        x = apply_repetition_penalty(penalty_ids, x, penalty)
    probs = softmax(top_k_mask(x, top_k) / temperature)
    out, margins = [], []
    for i in range(x.shape[0]):
        keep = nucleus(probs[i], top_p)
        pr = np.where(keep, probs[i], 0.0)
        cdf = np.cumsum(pr)
        u = float(uniform[i]) * cdf[-1]
        hit = np.nonzero(cdf > u)[0]
        out.append(int(hit[0]) if hit.size else int(np.nonzero(keep)[0][-1]))
        margins.append(float(np.abs(cdf[keep] - u).min()))
    return np.asarray(out, dtype=np.int64), np.asarray(margins)


def argmax_penalised(logits: np.ndarray, penalty_ids, penalty: float) -> np.ndarray:
    """static...:298-307 / dynamic...:265-274 for T < 0.05 with repetition_penalty > 1.01: arg-max (first
    occurrence) of the penalised logits."""
    return apply_repetition_penalty(penalty_ids, logits, penalty).argmax(axis=-1).astype(np.int64)


def exact_distribution(logits_row: np.ndarray, temperature: float, top_k: int, top_p: float) -> np.ndarray:
    """The distribution both engines sample from (up to the static engine's bounded rejection rounds): the
    top-p nucleus of softmax(top-k(logits) / T), renormalised.  Used to pin the algorithm against flashinfer's own
    samplers on the GPU box."""
    probs = softmax(top_k_mask(np.asarray(logits_row, dtype=np.float64)[None] / temperature, top_k))[0]
    keep = nucleus(probs, top_p)
    pr = np.where(keep, probs, 0.0)
    return pr / pr.sum()

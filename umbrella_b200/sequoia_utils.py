"""Sequoia growmaps: schema, loader, builders (reference umbrella/sequoia_utils.py:40-130 and the
tree set-up of static_speculation_engine.py:49-86).

Growmap JSON schema (kept byte-compatible): ``{roots, branches, Successors, mask, depth, size}``.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

DEFAULT_ACC = [0.65, 0.2, 0.1, 0.05]          # sequoia_utils.py:7
TREES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trees")


def successor_list_to_mask(successors: list[list[int]]) -> list[list[int]]:
    """mask[i][j] = 1 iff j is i or an ancestor of i (sequoia_utils.py:40-79)."""
    n = len(successors)
    parent = [-1] * n
    for v, succ in enumerate(successors):
        for c in succ:
            parent[c] = v
    mask = [[0] * n for _ in range(n)]
    for i in range(n):
        j = i
        while j >= 0:
            mask[i][j] = 1
            j = parent[j]
    return mask


def growmap_from_successors(successors: list[list[int]]) -> dict:
    n = len(successors)
    depth = [0] * n
    for v in range(n):
        for c in successors[v]:
            if c <= v:
                raise ValueError("growmaps are stored in level order (parent < child)")
            depth[c] = depth[v] + 1
    levels = max(depth) + 1
    roots = [[v for v in range(n) if depth[v] == d] for d in range(levels)]
    branches = [[len(successors[v]) for v in level] for level in roots]
    return {"roots": roots, "branches": branches, "Successors": [list(s) for s in successors],
            "mask": successor_list_to_mask(successors), "depth": depth, "size": n}


def generate_sequoia_tree(width: int, depth: int, acc: list[float] | None = None, json_file: str | None = None) -> dict:
    """Greedy tree constructor with the reference's semantics (sequoia_utils.py:83-130): per level,
    the ``width`` best (parent, rank) candidates by accumulated log-acceptance (float32, torch.topk
    tie order), children attached to the *sorted* parent list."""
    if acc is None:
        assert width <= 4, "Using default acceptance rate vector, require width<=4"
        acc = DEFAULT_ACC
    log_acc = torch.log(torch.tensor(acc, dtype=torch.float32))
    beams = len(acc)
    successors: list[list[int]] = [[]]
    level_scores = torch.zeros(1)
    level_start = 0
    for lvl in range(depth):
        cand = log_acc.repeat(level_scores.numel()) + level_scores.repeat_interleave(beams)
        best, idx = cand.topk(k=width)
        parents = sorted((idx // beams + level_start).tolist())
        first_child = len(successors)
        successors.extend([[] for _ in range(width)])
        for j, p in enumerate(parents):
            successors[p].append(first_child + j)
        level_scores = best
        level_start = first_child
    gm = growmap_from_successors(successors)
    if json_file is not None:
        with open(json_file, "w") as f:
            json.dump(gm, f)
    return gm


def generate_budget_tree(budget: int, max_depth: int, acc: list[float] | None = None, json_file: str | None = None) -> dict:
    """The ``budget`` nodes of highest path probability within ``max_depth`` levels below the root (best-first over
    (parent, rank) candidates; a node's probability is the product of ``acc[rank]`` along its path, so the set is
    closed under "parent of" and "elder sibling of" and maximises ``expected_accept_length`` for its size).

    Why a budget instead of the reference's fixed ``width`` per level (sequoia_utils.py:83-130): on a B200 the verify
    pass costs the same for every tree of <= 16 nodes -- the token rows are the padded N operand of the weight-streaming
    MMA and one 16-row attention slab per KV-head group -- and the draft side costs one forward per level, so the best
    tree is "as many nodes as the padding holds" at the depth ``tune_budget_tree`` picks from the measured costs."""
    import heapq
    acc = DEFAULT_ACC if acc is None else acc
    assert budget >= 1 and max_depth >= 0
    nodes = [(1.0, 0, -1, 0)]                               # (probability, depth, parent, rank) in pick order
    heap = [(-acc[0], 0, 1, 0, 0)] if max_depth >= 1 else []   # (-probability, tie-break, depth, parent, rank)
    tie = 1
    while len(nodes) < budget and heap:
        negp, _, d, par, r = heapq.heappop(heap)
        idx = len(nodes)
        nodes.append((-negp, d, par, r))
        if r + 1 < len(acc):
            heapq.heappush(heap, (-(nodes[par][0] * acc[r + 1]), tie, d, par, r + 1))
            tie += 1
        if d < max_depth:
            heapq.heappush(heap, (-(-negp * acc[0]), tie, d + 1, idx, 0))
            tie += 1
    # growmaps are stored in level order, children of a node contiguous and ordered by rank
    new_id, order = {0: 0}, [0]
    for d in range(1, max_depth + 1):
        level = sorted((i for i, n in enumerate(nodes) if n[1] == d), key=lambda i: (new_id[nodes[i][2]], nodes[i][3]))
        for i in level:
            new_id[i] = len(order)
            order.append(i)
    successors: list[list[int]] = [[] for _ in order]
    for i in order[1:]:
        successors[new_id[nodes[i][2]]].append(new_id[i])
    gm = growmap_from_successors(successors)
    if json_file is not None:
        with open(json_file, "w") as f:
            json.dump(gm, f)
    return gm


def tune_budget_tree(budget: int, fixed_ms: float, draft_forward_ms: float, acc: list[float] | None = None,
                     max_depth: int = 8, lazy_leaf: bool = True) -> tuple[dict, float]:
    """The budget tree whose depth maximises expected accepted tokens per second under a two-parameter cost model of one
    iteration: ``fixed_ms`` (target verify + accept + gathers: flat in the tree size up to the 16-row padding) plus one
    draft forward of ``draft_forward_ms`` per level below the root that is fed to the draft model (the leaf level is
    not, with the lazily computed leaf K/V).  Returns (growmap, predicted tokens per second).

    Measured on a B200 for the 70B-AWQ target + 1B draft: fixed 10.0 ms, 1.05 ms per draft forward ->
    depth 3, ``b200_sequoia_tree-16.json`` (packaged): 3.43 expected tokens per verify in 13.2 ms against 3.36 in 14.2 ms
    for ``sequoia_tree-3x4.json`` (13 nodes, depth 4)."""
    best, best_rate = None, -1.0
    for depth in range(1, max_depth + 1):
        gm = generate_budget_tree(budget, depth, acc)
        levels = max(gm["depth"])                                  # a small budget may not reach `depth`
        forwards = levels if lazy_leaf else levels + 1
        rate = 1e3 * expected_accept_length(gm, acc) / (fixed_ms + draft_forward_ms * max(forwards, 1))
        if rate > best_rate + 1e-12:
            best, best_rate = gm, rate
    return best, best_rate


def load_growmap(path: str) -> dict:
    """Load a growmap.  The Sequoia trees the reference ships (umbrella/trees/*.json) are packaged in compact form
    -- ``trees/growmaps.json``: file name -> successor lists -- and rebuilt into the full schema on load, so a bare
    file name or a path into the reference checkout (``../umbrella/trees/x.json``) both resolve by base name."""
    if os.path.exists(path):
        with open(path) as f:
            gm = json.load(f)
    else:
        with open(os.path.join(TREES_DIR, "growmaps.json")) as f:
            packaged = json.load(f)
        name = os.path.basename(path)
        if name not in packaged:
            raise FileNotFoundError(f"{path}: no such growmap file, and not one of the packaged trees {sorted(packaged)}")
        gm = growmap_from_successors(packaged[name])
    if "Successors" not in gm or "size" not in gm:
        raise ValueError(f"{path}: not a growmap")
    if "mask" not in gm or "depth" not in gm or "roots" not in gm or "branches" not in gm:
        gm = growmap_from_successors(gm["Successors"])
    return gm


def tree_bits_from_mask(mask) -> np.ndarray:
    """bool [N, N] ancestor-or-self mask -> int32 words [N, ceil(N/32)], bit j%32 of word j//32."""
    m = np.asarray(mask).astype(bool)
    n = m.shape[0]
    words = (m.shape[1] + 31) // 32
    padded = np.zeros((n, words * 32), dtype=np.uint64)
    padded[:, :m.shape[1]] = m
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    out = (padded.reshape(n, words, 32) * weights).sum(axis=-1).astype(np.uint32)
    return out.view(np.int32)


def expected_accept_length(gm: dict, acc: list[float] | None = None) -> float:
    """E[tokens per verify] when the rank-r child of any node is accepted with probability acc[r]
    (sum over nodes of the product of edge probabilities) -- the figure a scripted run converges to."""
    acc = DEFAULT_ACC if acc is None else acc
    succ = gm["Successors"]
    prob = [0.0] * gm["size"]
    prob[0] = 1.0
    for v in range(gm["size"]):
        for r, c in enumerate(succ[v]):
            prob[c] = prob[v] * (acc[r] if r < len(acc) else 0.0)
    return float(sum(prob))


def target_distribution(logits: torch.Tensor, temperature: float = 0.0, topp: float = 1.0) -> torch.Tensor:
    """What the verify step draws the committed token from, per row of ``logits`` [T, V]: one-hot arg-max for the
    greedy engines (temperature < 0.05, static_speculation_engine.py:298-300), else softmax(logits / T) restricted to
    the nucleus {j : mass strictly above p_j < topp} and renormalised (the set umb_sample_f32 keeps)."""
    logits = logits.float()
    if temperature < 0.05:
        return torch.nn.functional.one_hot(logits.argmax(-1), logits.shape[-1]).float()
    p = torch.softmax(logits / temperature, dim=-1)
    if topp < 1.0:
        sp, order = p.sort(dim=-1, descending=True)
        keep_sorted = (sp.cumsum(-1) - sp) < topp
        keep = torch.zeros_like(keep_sorted).scatter(-1, order, keep_sorted)
        p = torch.where(keep, p, torch.zeros_like(p))
        p = p / p.sum(-1, keepdim=True)
    return p


def rank_acceptance(draft_logits: torch.Tensor, target_logits: torch.Tensor, width: int, temperature: float = 0.0,
                    topp: float = 1.0) -> torch.Tensor:
    """Sum over the T rows of P(the token the target commits is the draft's rank-r candidate), r < width.  Greedy: the
    count of rows whose target arg-max equals the draft's rank-r token -- the accumulation of
    examples/construct_sequoia.py:79-83."""
    cand = draft_logits.float().topk(k=width, dim=-1).indices                  # [T, width]
    return target_distribution(target_logits, temperature, topp).gather(-1, cand).sum(dim=0)


@torch.inference_mode()
def measure_acceptance_rate(draft_model, target_model, dataset, temperature: float = 0.6, topp: float = 0.9,
                            repetition_penalty: float = 1.05, width: int = 4, max_length: int | None = None,
                            normalize: bool = True) -> list[float]:
    """Acceptance-rate vector for ``generate_sequoia_tree(acc=...)``: acc[r] = how often the target's next token is the
    draft's rank-r candidate, teacher-forced over ``dataset``.

    The reference declares this function and leaves it empty (sequoia_utils.py:133-135); the measurement it stands for is
    the loop of examples/construct_sequoia.py:58-90, restated here over the models' reference ``inference`` seam:

    * ``dataset``: iterable of ``(context_ids, continuation_ids)`` pairs (1-D int tensors / lists).  Both models run
      once over ``context + continuation`` with a causal mask; the rows that predict the continuation are compared.
    * ``temperature`` / ``topp`` select the verify rule the tree is built for (``target_distribution``);
      ``repetition_penalty`` > 1 is applied to the target's rows over the tokens that precede each row, as verify does.
    * ``normalize=False`` returns the raw sums, which is what the reference script passes on (construct_sequoia.py:90);
      rates (sums / rows) are the meaningful input to the tree builder: log-rates above 0 make every deeper node look
      better than any shallower one.

    Both models are ``clear()``-ed after every sample.  ``draft_model`` / ``target_model`` are model objects with the
    reference surface (``inference``, ``clear``, ``device``); the loading of names is ``AutoModelLM``'s business.
    """
    from .speculation.speculation_utils import apply_repetition_penalty, make_causal_mask
    dev = target_model.device
    sums = torch.zeros(width, dtype=torch.float64)
    rows_total = 0
    for context, continuation in dataset:
        ctx = torch.as_tensor(context, dtype=torch.long).reshape(-1)
        cont = torch.as_tensor(continuation, dtype=torch.long).reshape(-1)
        if cont.numel() == 0 or ctx.numel() == 0:
            continue
        ids = torch.cat([ctx, cont])
        if max_length is not None and ids.numel() > max_length:
            ids = ids[:max_length]
        T = ids.numel()
        n_pred = T - ctx.numel()                       # rows ctx-1 .. T-2 predict the continuation tokens
        if n_pred <= 0:
            continue
        ids_dev = ids.to(dev)[None]
        pos = torch.arange(T, device=dev)[None]
        mask = make_causal_mask((1, T), dev)
        store = torch.arange(T, device=dev)
        per_model = []
        for m in (target_model, draft_model):
            out = m.inference(input_ids=ids_dev, position_ids=pos, attention_mask=mask, storage_ids=store)
            per_model.append(out[0, T - n_pred - 1:T - 1].float().cpu())
            m.clear()
        tgt, drf = per_model
        if repetition_penalty > 1.0 and temperature >= 0.05:
            tgt = torch.stack([apply_repetition_penalty(ids[None, :T - n_pred + i], tgt[i:i + 1], repetition_penalty)[0]
                               for i in range(n_pred)])
        sums += rank_acceptance(drf, tgt, width, temperature, topp).double()
        rows_total += n_pred
    if normalize and rows_total:
        sums = sums / rows_total
    return sums.tolist()

"""Oracle: Sequoia growmaps, tree accept, KV gather, state advance (numpy, integer-exact).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Everything here is integer / boolean work and must be
bit-exact with the reference; ``tests/test_oracle_tree.py`` pins it against
fixtures produced by running the reference's own Python on CPU
(``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

# umbrella/sequoia_utils.py:7
DEFAULT_ACC = [0.65, 0.2, 0.1, 0.05]


# --------------------------------------------------------------------------------------
# growmap construction
# --------------------------------------------------------------------------------------
def successors_to_parents(successors: list[list[int]]) -> np.ndarray:
    """parents[child] = v for every child in Successors[v]; parents[0] = 0.

    Restates umbrella/speculation/static_speculation_engine.py:73-75 (int32 zeros, then
    ``parents[successor] = v``).
    """
    parents = np.zeros(len(successors), dtype=np.int32)
    for v, succ in enumerate(successors):
        for c in succ:
            parents[c] = v
    return parents


def successors_to_mask(successors: list[list[int]]) -> list[list[int]]:
    """mask[i][j] = 1 iff j is an ancestor of i or i itself.

    Same result as umbrella/sequoia_utils.py:40-79 (BFS over the reversed graph); here a
    parent walk, which is equivalent because every node has exactly one parent.
    """
    n = len(successors)
    parents = successors_to_parents(successors)
    has_parent = np.zeros(n, dtype=bool)
    for succ in successors:
        for c in succ:
            has_parent[c] = True
    mask = [[0] * n for _ in range(n)]
    for i in range(n):
        j = i
        mask[i][j] = 1
        while has_parent[j]:
            j = int(parents[j])
            mask[i][j] = 1
    return mask


def growmap_from_successors(successors: list[list[int]]) -> dict:
    """Rebuild the full growmap JSON schema {roots, branches, Successors, mask, depth, size}
    (umbrella/sequoia_utils.py:116-123) from the successor lists alone.

    ``roots[d]`` = node ids at depth d (level order, contiguous); ``branches[d][j]`` = number of
    children of the j-th node of level d.
    """
    n = len(successors)
    parents = successors_to_parents(successors)
    depth = [0] * n
    for v in range(n):  # parent < child holds for every shipped tree (level order)
        for c in successors[v]:
            assert c > v, "growmaps are stored in level order (parent < child)"
            depth[c] = depth[v] + 1
    max_d = max(depth)
    roots = [[v for v in range(n) if depth[v] == d] for d in range(max_d + 1)]
    branches = [[len(successors[v]) for v in level] for level in roots]
    return {
        "roots": roots,
        "branches": branches,
        "Successors": [list(s) for s in successors],
        "mask": successors_to_mask(successors),
        "depth": depth,
        "size": n,
    }


def generate_sequoia_tree(width: int, depth: int, acc: list[float] | None = None) -> dict:
    """Greedy Sequoia constructor, restating umbrella/sequoia_utils.py:83-130.

    Quirks kept on purpose (they decide the shape of the shipped trees):
      * the per-level score list is stored in *top-k order*, while children are attached to the
        *sorted* parent list (sequoia_utils.py:107-113) -- so ``score[i+1][c]`` is the c-th best
        candidate score, not the score of the c-th child;
      * arithmetic is float32 and candidate selection is ``torch.topk`` -- exact 3-way score ties
        occur with DEFAULT_ACC, so the tie order of that primitive is part of the result.
    """
    if acc is None:
        assert width <= 4, "Using default acceptance rate vector, require width<=4"
        acc = DEFAULT_ACC
    import torch  # torch.topk's tie order decides the shipped shapes (3-way ties occur with DEFAULT_ACC)
    acc_t = torch.log(torch.Tensor(acc))
    num_beams = len(acc_t)
    size = width * depth + 1
    root = [[0]]
    score = [[0]]
    successor: list[list[int]] = [[]]
    branches = [[0]]
    tree_depth = [0]
    for i in range(depth):
        root.append(list(range(i * width + 1, (i + 1) * width + 1)))
        branches.append([0] * width)
        tree_depth.extend([i + 1] * width)
        successor.extend([[] for _ in range(width)])
        current = torch.Tensor(score[i]).repeat_interleave(num_beams)
        cand = acc_t.repeat(1 if i == 0 else width) + current
        sel_score, sel_idx = cand.topk(k=width)
        score.append(sel_score)
        offset = 0 if i == 0 else (i - 1) * width + 1
        sel_parents = (sel_idx // num_beams + offset).tolist()
        for child, parent in enumerate(sorted(sel_parents)):
            successor[parent].append(child + i * width + 1)
            branches[i][parent - offset] += 1
    return {
        "roots": root,
        "branches": branches,
        "Successors": successor,
        "mask": successors_to_mask(successor),
        "depth": tree_depth,
        "size": size,
    }


def tree_bits_from_mask(mask: np.ndarray) -> np.ndarray:
    """Pack a boolean [N, N] ancestor mask into u32 words [N, ceil(N/32)], bit j%32 of word j//32.

    This is the compact operand of the product's attention/accept kernels; the dense form is the
    reference's (static_speculation_engine.py:66-68,79).
    """
    mask = np.asarray(mask).astype(bool)
    n = mask.shape[0]
    w = (n + 31) // 32
    out = np.zeros((n, w), dtype=np.uint32)
    for i in range(n):
        for j in range(mask.shape[1]):
            if mask[i, j]:
                out[i, j // 32] |= np.uint32(1) << np.uint32(j % 32)
    return out


# --------------------------------------------------------------------------------------
# attention-mask bookkeeping (reference keeps an O(S^2) master mask; the product keeps
# (prefix_len, tree_bits) -- these helpers prove the two agree)
# --------------------------------------------------------------------------------------
def make_causal_mask(n: int) -> np.ndarray:
    """umbrella/speculation/speculation_utils.py:4-14: mask[i, j] = (j <= i)."""
    idx = np.arange(n)
    return idx[None, :] < (idx + 1)[:, None]


def static_master_mask(max_length: int, tree_mask: np.ndarray) -> np.ndarray:
    """static_speculation_engine.py:55-57,79: bool [S, 2S]; [:S,:S] causal; last N x N of the
    left half = tree mask."""
    s = max_length
    n = tree_mask.shape[0]
    m = np.zeros((s, 2 * s), dtype=bool)
    m[:s, :s] = make_causal_mask(s)
    m[s - n:s, s - n:s] = tree_mask.astype(bool)
    return m


def static_mask_window(master: np.ndarray, max_length: int, num_nodes_this_iter: int) -> np.ndarray:
    """static_speculation_engine.py:150,345: rows [S-k, S), cols [S-k, 2S-k) with k = num_nodes+N."""
    s, k = max_length, num_nodes_this_iter
    return np.ascontiguousarray(master[s - k:s, s - k:2 * s - k])


def dynamic_master_mask(max_length: int, tree_size: int) -> np.ndarray:
    """dynamic_speculation_engine.py:53,59: bool [S, S]; causal on the first S-N+1 rows/cols."""
    s = max_length
    m = np.zeros((s, s), dtype=bool)
    c = s - tree_size + 1
    m[:c, :c] = make_causal_mask(c)
    return m


def dynamic_mask_window(master: np.ndarray, max_length: int, num_nodes_this_iter: int) -> np.ndarray:
    """dynamic_speculation_engine.py:105,320."""
    s, k = max_length, num_nodes_this_iter
    return np.ascontiguousarray(master[s - k:s, s - k:s])


def dense_rows_from_compact(prefix_len: int, tree_mask: np.ndarray, kv_len: int) -> np.ndarray:
    """Rows of the reference's ``attention_mask[num_nodes:num_nodes+N, :kv_len]`` rebuilt from the
    compact form: every tree row sees all of [0, prefix_len) plus its ancestors-or-self."""
    n = tree_mask.shape[0]
    out = np.zeros((n, kv_len), dtype=bool)
    out[:, :prefix_len] = True
    w = min(n, kv_len - prefix_len)
    out[:, prefix_len:prefix_len + w] = tree_mask[:, :w]
    return out


# --------------------------------------------------------------------------------------
# accept / verify bookkeeping
# --------------------------------------------------------------------------------------
def tree_accept(sampled: np.ndarray, speculated: np.ndarray, parents: np.ndarray,
                tree_mask: np.ndarray, node_in_path: np.ndarray):
    """The accept block, static_speculation_engine.py:312-325 == dynamic...:283-300.

    sampled[i]    target model's token after tree node i              (int64 [N])
    speculated[i] draft token sitting at tree node i                  (int64 [N])
    Returns (accept_path int64[len] relative node ids ascending, bonus token).
    """
    ref_tokens = sampled[parents]
    accept = ref_tokens == speculated
    accept[0] = True
    cnt = (accept[None, :] & tree_mask.astype(bool)).astype(np.int32).sum(axis=-1)
    accept_path = np.nonzero(cnt == node_in_path)[0].astype(np.int64)
    bonus = sampled[accept_path[-1]]
    return accept_path, bonus


def find_first_element_position(values, elements) -> int:
    """umbrella/speculation/speculation_utils.py:316-338."""
    for idx, v in enumerate(list(values)):
        if v in elements:
            return idx
    return -1


class VerifyState:
    """The engine-side integer state touched by verify (static_speculation_engine.py:51-56)."""

    def __init__(self, max_length: int, tree_depth: np.ndarray):
        self.max_length = max_length
        self.tokens = np.zeros(max_length, dtype=np.int64)
        self.position_ids = np.zeros(max_length, dtype=np.int64)
        self.depth = np.asarray(tree_depth, dtype=np.int64)
        self.tree_size = len(self.depth)
        self.num_nodes = 0

    def prefill(self, input_ids: np.ndarray, next_token: int):
        """static_speculation_engine.py:143-177 minus the model calls."""
        p = len(input_ids)
        self.num_nodes += p
        self.tokens[:p] = input_ids
        self.position_ids[:p] = np.arange(p)
        self.position_ids[p:p + self.tree_size] = p + self.depth
        self.tokens[self.num_nodes] = next_token

    def copy(self):
        o = VerifyState(self.max_length, self.depth)
        o.tokens = self.tokens.copy()
        o.position_ids = self.position_ids.copy()
        o.num_nodes = self.num_nodes
        return o


def verify_step(state: VerifyState, sampled: np.ndarray, parents: np.ndarray,
                tree_mask: np.ndarray, node_in_path: np.ndarray, eos_tokens: list[int]):
    """Everything in verify() after the target forward + sampling and before/after the KV gather:
    static_speculation_engine.py:312-349 (dynamic_speculation_engine.py:283-323 is the same
    sequence with ``node_in_path = depth + 1``).

    Mutates ``state`` exactly like the reference mutates ``tokens`` / ``position_ids`` /
    ``num_nodes``; returns (continue_generation, accept_path_abs int64[len], accept_length).
    """
    n, N = state.num_nodes, state.tree_size
    speculated = state.tokens[n:n + N].copy()
    accept_path, bonus = tree_accept(sampled, speculated, parents, tree_mask, node_in_path)
    accept_length = len(accept_path)
    state.tokens[n:n + accept_length] = speculated[accept_path]
    state.tokens[n + accept_length] = bonus

    cont = True
    eos_pos = find_first_element_position(state.tokens[n:n + accept_length + 1].tolist(), eos_tokens)
    if eos_pos >= 0:
        cont = False
        accept_path = accept_path[:eos_pos]
        accept_length = len(accept_path)
    accept_path_abs = accept_path + n

    state.num_nodes = n + accept_length
    if accept_length > 0:
        state.position_ids[n:state.num_nodes] = state.position_ids[accept_path_abs]
    state.position_ids[state.num_nodes:state.num_nodes + N] = state.num_nodes + state.depth
    return cont, accept_path_abs, accept_length


# --------------------------------------------------------------------------------------
# KV gather
# --------------------------------------------------------------------------------------
def kv_gather(cache: np.ndarray, indices: np.ndarray, offset: int, seq_axis: int = 1,
              zero_tail: bool = True) -> int:
    """umbrella/attn/cache.py:41-49 (seq_axis=1, layout [L,S,Hkv,D]) and :136-144 (seq_axis=2,
    layout [L,Hkv,S,D]).  RHS is materialised first (numpy fancy indexing copies), so overlapping
    source/destination rows are safe exactly as in the reference.  Returns the new kv_offset."""
    idx = np.asarray(indices, dtype=np.int64)
    m = len(idx)
    src = np.take(cache, idx, axis=seq_axis)
    sl = [slice(None)] * cache.ndim
    sl[seq_axis] = slice(offset, offset + m)
    cache[tuple(sl)] = src
    if zero_tail:
        sl[seq_axis] = slice(offset + m, None)
        cache[tuple(sl)] = 0
    return offset + m


# --------------------------------------------------------------------------------------
# scripted acceptance (benchmark input generator) -- not a reference function
# --------------------------------------------------------------------------------------
def scripted_samples(successors: list[list[int]], spec_tokens: np.ndarray, rng: np.random.Generator,
                     acc=DEFAULT_ACC, vocab_hi: int = 30000) -> np.ndarray:
    """Build a ``sampled`` vector that accepts, at each node, its rank-r child with probability
    acc[r] (else a token that matches no child)."""
    n = len(successors)
    sampled = np.zeros(n, dtype=np.int64)
    for v in range(n):
        kids = successors[v]
        u = rng.random()
        c = 0.0
        pick = None
        for r, child in enumerate(kids):
            if r < len(acc):
                c += acc[r]
                if u < c:
                    pick = child
                    break
        if pick is not None:
            sampled[v] = spec_tokens[pick]
        else:
            kid_tokens = {int(spec_tokens[k]) for k in kids}
            t = int(rng.integers(3, vocab_hi))
            while t in kid_tokens:
                t = int(rng.integers(3, vocab_hi))
            sampled[v] = t
    return sampled


# --------------------------------------------------------------------------------------
# dynamic (SpecExec-style) tree growth
# --------------------------------------------------------------------------------------
def dynamic_tree_level(top_logits: np.ndarray, token_ids: np.ndarray, score_history: np.ndarray, width: int):
    """One level of DynamicSpeculationEngine.build_tree after ``draft_logits.topk(num_beams)``
    (umbrella/speculation/dynamic_speculation_engine.py:235-245), fp32 as the reference computes it.

    top_logits / token_ids  [dec_len, num_beams]  the draft's top-``num_beams`` logits and their ids
    score_history           [dec_len]             tree_score of the level's nodes
    Returns (score[width] fp32, tokens[width] int64, parent_row[width] int64 = index into the level),
    best candidate first (torch.topk order; equal scores keep the lower flat index first).
    """
    x = np.asarray(top_logits, dtype=np.float32)
    e = np.exp(x - x.max(axis=-1, keepdims=True), dtype=np.float32)
    p = (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    cand = (np.asarray(score_history, dtype=np.float32)[:, None] + np.log(p + np.float32(1e-4), dtype=np.float32))
    cand = cand.astype(np.float32).reshape(-1)
    order = np.argsort(-cand, kind="stable")[:width]
    beams = x.shape[1]
    return cand[order], np.asarray(token_ids).reshape(-1)[order].astype(np.int64), (order // beams).astype(np.int64)


def dynamic_build_tree(draft_logits_fn, tokens: np.ndarray, num_nodes: int, width: int, num_beams: int, depth: int):
    """DynamicSpeculationEngine.build_tree (dynamic_speculation_engine.py:215-248) over a draft model given
    as ``draft_logits_fn(token_ids int64[q]) -> fp32 [q, V]`` (a context-free model is enough for the tree
    bookkeeping; the reference's forward also takes positions / mask / KV).

    Mutates ``tokens[num_nodes + 1 : num_nodes + N]``; returns (parents int32[N], tree_mask bool[N, N],
    tree_score fp32[N]) exactly as the reference leaves ``self.parents``, the tree block of
    ``attn_mask_this_iter`` and ``self.tree_score`` before verify().
    """
    N = width * depth + 1
    parents = np.zeros(N, dtype=np.int32)
    score = np.zeros(N, dtype=np.float32)
    mask = np.zeros((N, N), dtype=bool)
    mask[0, 0] = True
    first = 0
    for step in range(depth + 1):
        dec_len = width if step > 0 else 1
        if step < depth:
            logits = np.asarray(draft_logits_fn(tokens[num_nodes + first:num_nodes + first + dec_len]), dtype=np.float32)
            ids = np.argsort(-logits, axis=-1, kind="stable")[:, :num_beams]          # topk, sorted, ties: low id first
            top = np.take_along_axis(logits, ids, axis=-1)
            s, t, prow = dynamic_tree_level(top, ids, score[first:first + dec_len], width)
            new = first + dec_len
            score[new:new + width] = s
            tokens[num_nodes + new:num_nodes + new + width] = t
            parents[new:new + width] = prow + first
            mask[new:new + width] = mask[first + prow]                                  # copy the parent's row
            mask[np.arange(new, new + width), np.arange(new, new + width)] = True       # fill_diagonal_
        first += dec_len
    return parents, mask, score

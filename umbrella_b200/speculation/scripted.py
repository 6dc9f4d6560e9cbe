"""Scripted acceptance for benchmarks on synthetic weights.

Random-init draft and target models are uncorrelated, so "accepted tokens/sec" would degenerate to
one token per verify.  The script replaces the target's sampled token at every tree node, AFTER the
full target forward + arg-max has run (nothing is skipped in the timed region): with probability
``acc[r]`` the node's rank-r child is declared the target's choice (its draft token is copied), else
the target's own arg-max token is kept (which matches no child, up to chance).  The choices are
drawn up front from a seeded generator, so two implementations fed the same seed see the same
acceptance pattern.  ``acc`` defaults to the reference's DEFAULT_ACC (umbrella/sequoia_utils.py:7).
"""
from __future__ import annotations

import numpy as np
import torch

from ..sequoia_utils import DEFAULT_ACC


class ScriptedAcceptance:
    def __init__(self, successors, seed: int, max_iters: int, device, acc=None):
        acc = DEFAULT_ACC if acc is None else acc
        rng = np.random.default_rng(seed)
        n = len(successors)
        choice = np.full((max_iters, n), -1, dtype=np.int64)
        u = rng.random((max_iters, n))
        for v, kids in enumerate(successors):
            c = 0.0
            lo = np.zeros(max_iters)
            for r, child in enumerate(kids):
                if r >= len(acc):
                    break
                hi = lo + acc[r]
                choice[(u[:, v] >= lo) & (u[:, v] < hi), v] = child
                lo = hi
        self.choice = torch.from_numpy(choice).to(device)
        self.it = 0
        self.max_iters = max_iters
        self.it_dev = torch.zeros(1, dtype=torch.int64, device=device)     # iteration counter for graph replay

    def reset(self):
        self.it = 0
        self.it_dev.zero_()

    def snapshot(self):
        return int(self.it_dev.item())

    def restore(self, it: int):
        self.it_dev.fill_(it)

    def graph_call(self, engine, sampled: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
        """Capturable form: the iteration counter and the prefix length live on the device."""
        row = self.choice.index_select(0, self.it_dev % self.max_iters)[0]
        self.it_dev += 1
        idx = state[0:1].to(torch.int64) + row.clamp(min=0)
        picked = engine.tokens[0].gather(0, idx)
        self.last = torch.where(row >= 0, picked, sampled)
        return self.last

    def __call__(self, engine, sampled: torch.Tensor) -> torch.Tensor:
        row = self.choice[self.it % self.max_iters]
        self.it += 1
        n = engine.num_nodes
        tree_tokens = engine.tokens[0, n:n + row.numel()]
        picked = tree_tokens.gather(0, row.clamp(min=0))
        self.last = torch.where(row >= 0, picked, sampled)      # kept for tests / cross-implementation replay
        return self.last


class DynamicScriptedAcceptance:
    """The same script for the dynamic engine, whose tree is regrown every iteration: per (iteration, node) a seeded
    draw decides which child *rank* the target "chooses" (rank r with probability ``acc[r]``, none beyond); if the tree
    holds that child its token replaces the target's sampled token at the node, else the target's own token stays (and
    matches no child, up to chance).  A node's children are ranked by their cumulative draft score -- the tree keeps the
    best candidates of a level, so for one parent these are its top draft tokens in order, i.e. the in-tree rank is the
    draft rank.  Runs after the full target forward and sampling; device ops only, no host synchronisation."""

    def __init__(self, tree_size: int, seed: int, max_iters: int, device, acc=None):
        acc = DEFAULT_ACC if acc is None else acc
        rng = np.random.default_rng(seed)
        u = rng.random((max_iters, tree_size))
        edges = np.cumsum(np.asarray(acc, dtype=np.float64))
        want = np.searchsorted(edges, u, side="right")                   # rank whose interval holds u
        want[want >= len(acc)] = -1
        self.want = torch.from_numpy(want.astype(np.int64)).to(device)
        self.node = torch.arange(tree_size, device=device)
        self.it = 0
        self.max_iters = max_iters

    def reset(self):
        self.it = 0

    def child_ranks(self, parents: torch.Tensor, score: torch.Tensor) -> torch.Tensor:
        """rank[w] = how many siblings of w (same parent, root excluded) lie ahead of it by score (ties: lower index)."""
        node = self.node
        same = parents[:, None] == parents[None, :]
        ahead = (score[None, :] > score[:, None]) | ((score[None, :] == score[:, None]) & (node[None, :] < node[:, None]))
        return (same & ahead & (node[None, :] >= 1)).sum(dim=1)

    def __call__(self, engine, sampled: torch.Tensor) -> torch.Tensor:
        N, n = engine.tree_size, engine.num_nodes
        parents = engine.parents.to(torch.int64)
        want = self.want[self.it % self.max_iters]
        self.it += 1
        match = (self.child_ranks(parents, engine.tree_score) == want[parents]) & (self.node >= 1)
        slot = torch.where(match, parents, torch.full_like(parents, N))          # non-chosen nodes write to a spare slot
        out = torch.cat([sampled, sampled.new_zeros(1)]).scatter(0, slot, engine.tokens[0, n:n + N])
        self.last = out[:N]
        return self.last

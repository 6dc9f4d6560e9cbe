"""One replica per GPU: request stream i -> rank i % world, no data-path collective.

A request's speculative-decoding loop is a strictly sequential chain over private state, and the
reference never shards inside a request, so N GPUs run N independent engines.  ``torch.distributed``
is plumbing here: a barrier around the timed region and the reduction of the per-rank timings
(MAX of times, SUM of token counts) -- the same code under ``nccl`` on the GPU box and ``gloo`` in
the CPU tests.
"""
from __future__ import annotations

import os


def rank_env() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) as torchrun exports them; (0, 1, 0) for a plain launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def replica_seed(base: int, rank: int) -> int:
    """Independent request streams: every replica draws its own prompt / acceptance script."""
    return base + rank


def assign_requests(num_requests: int, world: int) -> list[list[int]]:
    """Round-robin partition of request ids over replicas (the multi-replica API front of 8f)."""
    out = [[] for _ in range(world)]
    for r in range(num_requests):
        out[r % world].append(r)
    return out


def reduce_over_replicas(times: list[float], counts: list[float], device=None):
    """Whole-job view of per-rank measurements: element-wise MAX of ``times`` and SUM of ``counts``
    over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(map(float, times)), list(map(float, counts))
    t = torch.tensor(times, dtype=torch.float64, device=device)
    c = torch.tensor(counts, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


def gather_over_replicas(values: list[float], device=None) -> list[list[float]]:
    """Every replica's ``values`` (one list per rank, rank order) -- the per-replica min / max SURVEY.md 8e asks for."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(map(float, values))]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def aggregate_throughput(times_ms: list[float], counts: list[float], device=None) -> float:
    """Units all ranks processed / the slowest rank's time (units per second)."""
    t, c = reduce_over_replicas(times_ms, counts, device)
    return c[0] / (t[0] * 1e-3)

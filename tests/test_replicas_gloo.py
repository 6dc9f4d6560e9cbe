"""Multi-replica plumbing on CPU: world_size-2 gloo processes run the launcher-side logic of bench.py
(per-rank seeds, barrier, MAX-of-times / SUM-of-counts reduction, rank-0-only reporting)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umbrella_b200 import replicas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    assert replicas.rank_env() == (rank, world, rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r "measured" (10 + 5 r) ms for (100 + r) tokens, e2e (2 + r) s for 50 tokens
    times, counts = replicas.reduce_over_replicas([10.0 + 5 * rank, 2.0 + rank], [100.0 + rank, 50.0])
    tput = replicas.aggregate_throughput([10.0 + 5 * rank], [100.0 + rank])
    per = replicas.gather_over_replicas([7.0 + rank, 1.0])
    assert per == [[7.0, 1.0], [8.0, 1.0]]                      # per-replica view (min / max in bench.py)
    dist.barrier()
    g = torch.Generator().manual_seed(replicas.replica_seed(7, rank))
    prompt = torch.randint(3, 30000, (8,), generator=g)
    gathered = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, prompt)
    if rank == 0:
        out.put((times, counts, tput, [t.tolist() for t in gathered]))
    dist.destroy_process_group()


def test_two_replicas_reduce_like_bench():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    times, counts, tput, prompts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert times == [15.0, 3.0]                  # slowest replica bounds the job
    assert counts == [201.0, 100.0]              # tokens add up over replicas
    assert abs(tput - 201.0 / 15e-3) < 1e-6
    assert prompts[0] != prompts[1]              # independent request streams


def test_single_process_is_identity():
    assert replicas.reduce_over_replicas([3.0], [4.0]) == ([3.0], [4.0])
    assert replicas.gather_over_replicas([3.0, 4.0]) == [[3.0, 4.0]]
    assert replicas.assign_requests(5, 2) == [[0, 2, 4], [1, 3]]
    assert replicas.assign_requests(3, 8)[:3] == [[0], [1], [2]]

"""Where one speculative-decoding iteration of the bench workload goes (CUDA events, eager launches vs the
whole-iteration graph): draft levels, target verify forward, accept + gathers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from umbrella_b200 import ops  # noqa: E402
from umbrella_b200.speculation.scripted import ScriptedAcceptance  # noqa: E402

dev = "cuda:0"
WL = bench.WORKLOADS[os.environ.get("WORKLOAD", "static70b")]
eng = bench.build_engine(dev, WL)
eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=1, max_iters=4096, device=dev)
prompt = torch.randint(3, 30000, (1, WL["prompt"]), device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps):
    s, e = ev(), ev()
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for graph in (False, True):
    eng.graph_iteration = graph
    eng.reset()
    eng._prefill(prompt)
    for _ in range(4):
        eng.build_tree(); eng.verify()
    it = timed(lambda: (eng.build_tree(), eng.verify()), 24)
    print(f"graph={graph}: iteration {it:.3f} ms")

eng.graph_iteration = False
eng.reset(); eng._prefill(prompt)
n, N = eng.num_nodes, eng.tree_size
dm, tm = eng.draft_model, eng.target_model


def draft_levels():
    dm.kv_cache.kv_offset = n
    for step in range(eng.tree_depth):
        a, b = eng.level_ranges[step]
        last = step == eng.tree_depth - 1
        lg = dm.verify_forward(eng.tokens[0, n + a:n + b], eng.position_ids[0, n + a:n + b], n, eng.tree_bits[a:b],
                               want_logits=not last)
        if not last:
            width = max(eng.branch_lists[step])
            lg.topk(k=width).indices.flatten().index_select(0, eng.sample_gather_indices[step])


def draft_forwards_only():
    dm.kv_cache.kv_offset = n
    for step in range(eng.tree_depth):
        a, b = eng.level_ranges[step]
        dm.verify_forward(eng.tokens[0, n + a:n + b], eng.position_ids[0, n + a:n + b], n, eng.tree_bits[a:b],
                          want_logits=step != eng.tree_depth - 1)


def target():
    tm.kv_cache.kv_offset = n
    tm.verify_forward(eng.tokens[0, n:n + N], eng.position_ids[0, n:n + N], n, eng.tree_bits, argmax_keys=eng._argmax_keys)


def tail():
    sampled = ops.argmax_keys_to_ids(eng._argmax_keys[:N])
    ops.tree_accept(sampled, eng.tokens[0].clone(), eng.position_ids[0].clone(), eng.parents, eng.depth, n, None, eng._result)
    dm.kv_cache.gather_from_result(eng._result, n, 0)
    tm.kv_cache.gather_from_result(eng._result, n, 0)


for name, fn in (("draft levels (5 forwards + top-k)", draft_levels), ("draft forwards only", draft_forwards_only),
                 ("target verify forward", target), ("argmax decode + accept + 2 gathers", tail)):
    fn(); fn()
    print(f"{name:40s} {timed(fn, 10):8.3f} ms (eager launches)")

# the same pieces as CUDA graphs (no host launch gaps)
for name, fn in (("draft levels", draft_levels), ("target verify forward", target)):
    g = torch.cuda.CUDAGraph()
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    print(f"{name:40s} {timed(g.replay, 10):8.3f} ms (graph replay)")

"""Stress: repeated reset + prefill (+ a few decode iterations) of the bench engine; reports the first failure."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from umbrella_b200.speculation.scripted import ScriptedAcceptance
dev = "cuda:0"
eng = bench.build_engine(dev, bench.WORKLOADS["static70b"])
eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=1, max_iters=4096, device=dev)
eng.graph_iteration = os.environ.get("GRAPH", "1") == "1"
prompt = torch.randint(3, 30000, (1, int(os.environ.get("PROMPT", "512"))), device=dev)
n_ok = 0
t0 = time.time()
for it in range(int(os.environ.get("REPS", "40"))):
    try:
        eng.reset()
        eng._prefill(prompt)
        torch.cuda.synchronize()
        for _ in range(int(os.environ.get("DECODE", "3"))):
            eng.build_tree(); eng.verify()
        torch.cuda.synchronize()
        n_ok += 1
    except Exception as e:
        print(f"FAILED at rep {it}: {str(e)[:200]}", flush=True)
        break
print(f"stress done: {n_ok} ok reps in {time.time() - t0:.1f} s; env PDL={os.environ.get('UMB_PDL', '1')} GRAPH={os.environ.get('GRAPH', '1')}")

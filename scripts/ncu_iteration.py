"""Driver for the per-kernel launch list: a few speculative-decoding iterations of the bench workload
(70B-AWQ target + 1B draft, 3x4 tree) after a short prompt, eager mode so that every kernel is a
stream launch ncu lists in order."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from umbrella_b200.speculation.scripted import ScriptedAcceptance
eng = bench.build_engine("cuda:0", 0)
eng.graph_iteration = os.environ.get("GRAPH", "0") == "1"
eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=1, max_iters=64, device="cuda:0")
prompt = torch.randint(3, 30000, (1, int(os.environ.get("PROMPT", "64"))), device="cuda:0")
eng._prefill(prompt)
torch.cuda.synchronize()
print("prefill done", flush=True)
for _ in range(int(os.environ.get("ITERS", "3"))):
    eng.build_tree(); eng.verify()
torch.cuda.synchronize()
print("done")

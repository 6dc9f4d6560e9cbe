"""Dynamic (SpecExec) engine at full scale on one B200 -- BASELINE.json config 3 shapes:
Llama-3.1-70B-AWQ target + Llama-3.1-8B-AWQ draft, tree width x depth (+1) nodes, synthetic weights.
Reports time per phase -- build_tree (depth + 1 draft forwards + tree growth kernels) and verify (chunked target
forward over N rows + accept + gathers) -- and accepted tokens/s under scripted acceptance: random weights make draft
and target uncorrelated (accept length ~1), so per tree node the target's choice is scripted to be the node's rank-r
child with probability DEFAULT_ACC[r] (speculation/scripted.py: DynamicScriptedAcceptance; SCRIPT=0 turns it off).  The
tree is still grown from the synthetic draft's scores, so its shape is that of a near-uniform draft distribution.
WIDTH / DEPTH / BEAMS / PROMPT / ITERS / SCRIPT from the environment."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200.models import llama as L  # noqa: E402
from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine  # noqa: E402
from umbrella_b200.speculation.scripted import DynamicScriptedAcceptance  # noqa: E402

dev = "cuda:0"
W, D, B = int(os.environ.get("WIDTH", "16")), int(os.environ.get("DEPTH", "24")), int(os.environ.get("BEAMS", "24"))
P, ITERS, S = int(os.environ.get("PROMPT", "512")), int(os.environ.get("ITERS", "12")), 8192
SCRIPT = os.environ.get("SCRIPT", "1") != "0"
eng = DynamicSpeculationEngine("draft", "target", device=dev, max_length=S, width=W, depth=D, num_beams=B)
t0 = time.time()
target = L.SyntheticLlama(L.LLAMA31_70B, max_length=S, awq=True, seed=1234, device=dev, gen_device=dev)
draft = L.SyntheticLlama(L.LLAMA31_8B, max_length=S, awq=True, seed=4321, device=dev, gen_device=dev)
eng.initialize(draft_model=draft, target_model=target)
eng._set_eos([])
print(f"models built in {time.time() - t0:.1f} s; tree {eng.tree_size} nodes", flush=True)
if SCRIPT:
    eng.acceptance_script = DynamicScriptedAcceptance(eng.tree_size, seed=100, max_iters=1024, device=dev)
prompt = torch.randint(3, 30000, (1, P), device=dev)
assert eng._prefill(prompt)


def ev():
    return torch.cuda.Event(enable_timing=True)


bt, vt, acc = [], [], []
for it in range(ITERS + 1):
    a, b, c = ev(), ev(), ev()
    before = eng.num_nodes
    a.record()
    eng.build_tree()
    b.record()
    eng.verify()
    c.record()
    torch.cuda.synchronize()
    if it:                                  # first iteration = warm-up
        bt.append(a.elapsed_time(b)); vt.append(b.elapsed_time(c)); acc.append(eng.num_nodes - before)
N = eng.tree_size
wbytes = target.weight_bytes()
chunks = (N + 127) // 128
line = {"workload": f"70B-AWQ target + 8B-AWQ draft, dynamic tree {W}x{D}+1 = {N} nodes, beams {B}, prompt {P}",
        "build_tree_ms": sum(bt) / len(bt), "verify_ms": sum(vt) / len(vt), "mean_accept_len": sum(acc) / len(acc),
        "scripted_acceptance": "DEFAULT_ACC per tree node (child rank by draft score)" if SCRIPT else None,
        "iterations": len(acc), "accepted_tokens_per_s": sum(acc) / ((sum(bt) + sum(vt)) * 1e-3),
        "draft_forwards": D + 1, "target_chunks": chunks,
        "verify_weight_passes_gbs": chunks * wbytes / (sum(vt) / len(vt) * 1e-3) / 1e9,
        "verify_tflops": 2 * N * (wbytes / 0.5195 if False else 70.55e9) / (sum(vt) / len(vt) * 1e-3) / 1e12}
print(json.dumps(line))

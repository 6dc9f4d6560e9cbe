"""Dynamic (SpecExec) engine at full scale on one B200 -- BASELINE.json config 3 shapes:
Llama-3.1-70B-AWQ target + Llama-3.1-8B-AWQ draft, tree width x depth (+1) nodes, synthetic weights.
Random weights make draft and target uncorrelated (accept length ~1), so this reports time per phase, not
tokens/s: build_tree (depth + 1 draft forwards + tree growth kernels) and verify (chunked target forward over
N rows + accept + gathers).   WIDTH / DEPTH / BEAMS / PROMPT / ITERS from the environment."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200.models import llama as L  # noqa: E402
from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine  # noqa: E402

dev = "cuda:0"
W, D, B = int(os.environ.get("WIDTH", "16")), int(os.environ.get("DEPTH", "24")), int(os.environ.get("BEAMS", "24"))
P, ITERS, S = int(os.environ.get("PROMPT", "512")), int(os.environ.get("ITERS", "4")), 8192
eng = DynamicSpeculationEngine("draft", "target", device=dev, max_length=S, width=W, depth=D, num_beams=B)
t0 = time.time()
target = L.SyntheticLlama(L.LLAMA31_70B, max_length=S, awq=True, seed=1234, device=dev, gen_device=dev)
draft = L.SyntheticLlama(L.LLAMA31_8B, max_length=S, awq=True, seed=4321, device=dev, gen_device=dev)
eng.initialize(draft_model=draft, target_model=target)
eng._set_eos([])
print(f"models built in {time.time() - t0:.1f} s; tree {eng.tree_size} nodes", flush=True)
prompt = torch.randint(3, 30000, (1, P), device=dev)
assert eng._prefill(prompt)


def ev():
    return torch.cuda.Event(enable_timing=True)


bt, vt, acc = [], [], []
for it in range(ITERS + 1):
    a, b, c = ev(), ev(), ev()
    a.record()
    eng.build_tree()
    b.record()
    eng.verify()
    c.record()
    torch.cuda.synchronize()
    if it:                                  # first iteration = warm-up
        bt.append(a.elapsed_time(b)); vt.append(b.elapsed_time(c)); acc.append(eng.last_accept_length)
N = eng.tree_size
wbytes = target.weight_bytes()
chunks = (N + 127) // 128
line = {"workload": f"70B-AWQ target + 8B-AWQ draft, dynamic tree {W}x{D}+1 = {N} nodes, beams {B}, prompt {P}",
        "build_tree_ms": sum(bt) / len(bt), "verify_ms": sum(vt) / len(vt), "mean_accept_len": sum(acc) / len(acc),
        "draft_forwards": D + 1, "target_chunks": chunks,
        "verify_weight_passes_gbs": chunks * wbytes / (sum(vt) / len(vt) * 1e-3) / 1e9,
        "verify_tflops": 2 * N * (wbytes / 0.5195 if False else 70.55e9) / (sum(vt) / len(vt) * 1e-3) / 1e12}
print(json.dumps(line))

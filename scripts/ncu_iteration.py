"""Driver for the per-kernel launch list (ncu --profile-from-start off): a few speculative-decoding
iterations of the bench workload (70B-AWQ-shaped target + 1B draft, 3x4 tree), eager mode so that every
kernel is a stream launch ncu lists in order.  Only the iterations between cudaProfilerStart/Stop are
profiled (the synthetic-weight generation is not).  LAYERS trims the target's depth (same layer shapes)
to keep the capture short; the per-layer kernel sequence is what is being listed.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/ncu_iteration.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from umbrella_b200.models import llama as L  # noqa: E402
from umbrella_b200.speculation.scripted import ScriptedAcceptance  # noqa: E402
from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine  # noqa: E402

dev = "cuda:0"
layers = int(os.environ.get("LAYERS", "4"))
WL = bench.WORKLOADS["static70b"]
eng = StaticSpeculationEngine("draft", "target", device=dev, growmap_path=WL["tree"], max_length=WL["max_length"],
                              temperature=0.0)
target = L.SyntheticLlama(L.LLAMA31_70B, max_length=WL["max_length"], awq=True, seed=1234, device=dev, gen_device=dev,
                          num_layers=layers)
draft = L.SyntheticLlama(L.LLAMA32_1B, max_length=WL["max_length"], awq=False, seed=4321, device=dev, gen_device=dev)
eng.initialize(draft_model=draft, target_model=target)
eng._set_eos([])
eng.graph_iteration = os.environ.get("GRAPH", "0") == "1"
eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=1, max_iters=64, device=dev)
prompt = torch.randint(3, 30000, (1, int(os.environ.get("PROMPT", "512"))), device=dev)
eng._prefill(prompt)
eng.build_tree()
eng.verify()          # warm-up iteration (module load, attribute setting)
torch.cuda.synchronize()
print("prefill done", flush=True)
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("ITERS", "2"))):
    eng.build_tree()
    eng.verify()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

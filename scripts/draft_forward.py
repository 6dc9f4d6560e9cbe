"""The draft model's per-level forward alone (Llama-3.2-1B fp16, prefix 512, ROWS token rows): CUDA-event time of a
CUDA graph holding 8 forwards (what the iteration graph replays), and -- under ncu --profile-from-start off -- the
in-order kernel list of two eager forwards.

    python scripts/draft_forward.py
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/draft_forward.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200.models import llama as L  # noqa: E402

dev = "cuda:0"
rows = int(os.environ.get("ROWS", "4"))
P = 512
draft = L.SyntheticLlama(L.LLAMA32_1B, max_length=2048, awq=False, seed=4321, device=dev, gen_device=dev)
prompt = torch.randint(3, 30000, (P,), device=dev)
draft.prefill_forward(prompt, torch.arange(P, device=dev), want_logits=False)
ids = torch.randint(3, 30000, (rows,), device=dev)
pos = torch.arange(P, P + rows, device=dev)
bits = torch.tril(torch.ones(rows, rows, dtype=torch.int32, device=dev))
bits = (bits << torch.arange(rows, device=dev, dtype=torch.int32)[None]).sum(-1, keepdim=True).to(torch.int32)


def fwd():
    draft.kv_cache.kv_offset = P
    return draft.verify_forward(ids, pos, P, bits)


for _ in range(3):
    fwd()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fwd()
fwd()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(8):
        fwd()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    g.replay()
e.record()
torch.cuda.synchronize()
print(f"draft forward, {rows} rows, prefix {P}: {s.elapsed_time(e) / 80 * 1e3:.1f} us per forward (graph of 8, UMB_SKINNY={os.environ.get('UMB_SKINNY', '1')})")

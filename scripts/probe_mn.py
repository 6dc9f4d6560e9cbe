"""tcgen05.mma with an MN-major B operand (V rows of a KV cache): which (LBO, SBO) reading is right."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops
g = torch.Generator().manual_seed(0)
a = torch.randn(128, 64, generator=g).to(torch.float16).cuda()
bt = torch.randn(64, 128, generator=g).to(torch.float16).cuda()
ref = a.float() @ bt.float()
for mode in (2, 3):
    d = ops.probe_umma(a, bt, mode)
    torch.cuda.synchronize()
    print(f"mode {mode}: max abs err {(d - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.2f})")

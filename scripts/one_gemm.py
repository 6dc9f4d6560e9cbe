"""One AWQ GEMM launch against the SIMT cross-check (debug helper: run under compute-sanitizer / with UMB_LIB_PATH)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops  # noqa: E402

M, K, N = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (13, 512, 256)))
dev = "cuda:0"
torch.manual_seed(0)
nbytes = ops.awq_packed_bytes(K, N)
packed = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev)
# sane fp16 scales inside every record (bytes 8192..8447): 0.01
rec = packed.view(-1, 8512)
rec[:, 8192:8448] = torch.full((128,), 0.01, dtype=torch.float16, device=dev).view(torch.uint8)
x = torch.randn(M, K, device=dev).to(torch.float16)
ws = ops.gemm_workspace(M, K, N, dev)
y = ops.awq_gemm(x, packed, K, N, workspace=ws)
torch.cuda.synchronize()
w = ops.awq_unpack(packed, K, N).float()
ref = x.float() @ w.t()
err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
print(f"M={M} K={K} N={N}: rel err {err:.2e}")

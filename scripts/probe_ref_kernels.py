"""GPU-box probe: can the reference's third-party kernels be called in this image, and how long does the first
call take (flashinfer JIT / pre-built modules from baseline/fi_ws)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(1, os.path.join(ROOT, "tests"))
t0 = time.time()
import torch
import refkernels
print("refkernels.available():", refkernels.available(), f"({time.time()-t0:.1f} s)")
dev = "cuda:0"
def step(name, fn):
    t = time.time()
    try:
        out = fn(); torch.cuda.synchronize()
        print(f"OK   {name}: {time.time()-t:.1f} s", tuple(out.shape) if hasattr(out, "shape") else out)
    except Exception as e:
        print(f"FAIL {name}: {time.time()-t:.1f} s {e!r}"[:600])
x = torch.randn(13, 8192, device=dev).half()
w = torch.ones(8192, device=dev).half()
step("flashinfer.rmsnorm", lambda: refkernels.rmsnorm(x, w, 1e-5))
q = torch.randn(13, 64, 128, device=dev).half(); k = torch.randn(77, 8, 128, device=dev).half(); v = torch.randn(77, 8, 128, device=dev).half()
mask = torch.ones(13, 77, dtype=torch.bool, device=dev)
step("flashinfer.single_prefill(custom_mask) D=128", lambda: refkernels.attention(q, k, v, mask))
step("flashinfer.single_prefill(custom_mask) D=64", lambda: refkernels.attention(q[..., :64].contiguous(), k[..., :64].contiguous(), v[..., :64].contiguous(), mask))
import flashinfer
logits = torch.randn(13, 128256, device=dev)
step("flashinfer.sampling.top_k_top_p_sampling_from_logits", lambda: flashinfer.sampling.top_k_top_p_sampling_from_logits(logits / 0.6, 32, 0.9))
step("flashinfer.sampling.top_p_renorm_probs", lambda: flashinfer.sampling.top_p_renorm_probs(torch.softmax(logits, -1), 0.9))
from oracle import awq as oawq
import numpy as np
qw, qz, sc = oawq.random_awq(8192, 1024, 128, np.random.default_rng(0))
trip = tuple(torch.from_numpy(a).to(dev) for a in (qw, qz, sc))
step("vllm awq_gemm", lambda: refkernels.awq_linear(x, trip))
print(f"total {time.time()-t0:.1f} s")

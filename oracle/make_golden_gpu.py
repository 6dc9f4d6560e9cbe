"""Golden vectors for the AWQ arithmetic from the UPSTREAM KERNEL, generated on a GPU box.

TEST INFRASTRUCTURE ONLY.  The reference calls ``awq_ext.gemm_forward_cuda`` / ``dequantize_weights_cuda``
(umbrella/quantization/awq_utils.py:67-77) from autoawq-kernels==0.0.8, which is not in this image.  vLLM 0.22
ships a build of the same upstream CUDA sources (device symbols ``vllm::awq::gemm_forward_4bit_cuda_m16nXk32``,
``vllm::awq::dequantize_weights``; python entry points ``vllm._custom_ops.awq_gemm / awq_dequantize``, same
argument order and ``split_k_iters = 8`` as the reference's call).  This script runs those kernels on seeded AWQ
triplets and writes ``gpurun_out/awq_upstream.npz``; the file is committed as ``tests/golden/awq_upstream.npz``:

    gpurun -- 'python oracle/make_golden_gpu.py'  &&  cp gpurun_out/awq_upstream.npz tests/golden/
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import awq as oawq  # noqa: E402


def main():
    os.environ.setdefault("VLLM_USE_TRITON_AWQ", "0")
    from vllm import _custom_ops as vops
    assert hasattr(torch.ops._C, "awq_gemm") and hasattr(torch.ops._C, "awq_dequantize"), "vLLM _C AWQ ops missing"
    dev = "cuda:0"
    out = {}
    for ci, (m, k, n, seed) in enumerate([(1, 256, 128, 0), (13, 512, 256, 1), (31, 1024, 384, 2), (16, 384, 512, 3)]):
        rng = np.random.default_rng(1000 + seed)
        qweight, qzeros, scales = oawq.random_awq(k, n, 128, rng)
        x = (rng.standard_normal((m, k)) * 0.5).astype(np.float16)
        tq, tz = torch.from_numpy(qweight).to(dev), torch.from_numpy(qzeros).to(dev)
        ts, tx = torch.from_numpy(scales).to(dev), torch.from_numpy(x).to(dev)
        deq = vops.awq_dequantize(tq, ts, tz, 0, 0, 0)                       # fp16 [K, N]
        y = vops.awq_gemm(tx, tq, ts, tz, 8)                                 # fp16 [M, N], split_k_iters = 8
        torch.cuda.synchronize()
        pre = "c%d_" % ci
        out[pre + "meta"] = np.array([m, k, n, 128], dtype=np.int64)
        out[pre + "qweight"], out[pre + "qzeros"], out[pre + "scales"], out[pre + "x"] = qweight, qzeros, scales, x
        out[pre + "dequant"] = deq.cpu().numpy()
        out[pre + "y"] = y.cpu().numpy()
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/awq_upstream.npz", **out)
    print("wrote gpurun_out/awq_upstream.npz", len(out), "arrays")


if __name__ == "__main__":
    main()

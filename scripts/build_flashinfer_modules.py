"""Pre-build the flashinfer JIT modules the reference-kernel comparator (tests/refkernels.py) and the reference
arm of bench.py call, for sm_100a, into baseline/fi_ws (git-ignored, travels to the GPU box with the snapshot).
nvcc cross-compiles them without a GPU.  Run once per container:

    python scripts/build_flashinfer_modules.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FLASHINFER_WORKSPACE_BASE"] = os.path.join(ROOT, "baseline", "fi_ws")
os.environ["FLASHINFER_CUDA_ARCH_LIST"] = "10.0a"

import torch  # noqa: E402


def main():
    from flashinfer.norm import gen_norm_module
    from flashinfer.prefill import gen_single_prefill_module
    from flashinfer.sampling import gen_sampling_module
    specs = [gen_norm_module(), gen_sampling_module()]
    for d in (128, 64):
        # fp16 q/k/v/o, no positional encoding, no sliding window, no soft cap, fp32 QK reduction (the fp16-reduction
        # variant the reference asks for does not compile in this image: prefill.cuh:182 needs boost_math)
        specs.append(gen_single_prefill_module("fa2", torch.float16, torch.float16, torch.float16, d, d, 0, False, False, False))
    for s in specs:
        t = time.time()
        s.build(verbose=False)
        print(f"built {s.name[:60]} in {time.time() - t:.0f} s")


if __name__ == "__main__":
    sys.exit(main())

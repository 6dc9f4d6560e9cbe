"""Build libumbrella_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m umbrella_b200.build            # incremental
    python -m umbrella_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libumbrella_b200.so")
SOURCES = ["api.cu", "tree.cu", "glue.cu", "attn.cu", "gemm.cu", "forward.cu", "sampler.cu"]
HEADERS = ["common.cuh", "internal.h", os.path.join("..", "..", "include", "umbrella_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libumbrella_b200.so cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None, variant: str = "",
          gemm_source: str | None = None) -> str:
    """``variant`` + ``extra_flags`` (e.g. ["-DUMB_DQ_GROUPS=4"]) build an experiment library
    libumbrella_b200.<variant>.so next to the product one; select it with UMB_LIB_PATH.  ``gemm_source`` swaps
    csrc/gemm.cu for another translation unit exporting the same entry points (csrc/experimental/gemm_ring.cu: the
    round-2 weight-request-ring / drain-form kernel, kept for the experiments recorded in profiles/r2_gemm_experiments.txt)."""
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + list(extra_flags or [])
    obj_dir = os.path.join(HERE, "build" + (("_" + variant) if variant else ""))
    LIB = os.path.join(HERE, f"libumbrella_b200.{variant}.so") if variant else globals()["LIB"]
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:
        src = os.path.join(CSRC, gemm_source if (gemm_source and s == "gemm.cu") else s)
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + flags + ["-I", CSRC, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    # python -m umbrella_b200.build [--force] [--variant NAME -DFOO=1 ...]
    argv = sys.argv[1:]
    variant = argv[argv.index("--variant") + 1] if "--variant" in argv else ""
    gemm = argv[argv.index("--gemm") + 1] if "--gemm" in argv else None
    print(build(force="--force" in argv, verbose=True, extra_flags=[a for a in argv if a.startswith("-D")],
                variant=variant, gemm_source=gemm))

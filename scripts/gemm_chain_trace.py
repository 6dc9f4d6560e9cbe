"""Timeline of ONE launch in the middle of a back-to-back chain of AWQ GEMM launches (trace build,
UMB_GEMM_TRACE_LAUNCH): where the fixed per-launch microseconds of the small GEMMs go.  Stamps are per CTA, printed
relative to the moment the stream predecessor completed as seen by this launch (dep_ok, griddepcontrol.wait returns)."""
import ctypes as C, os, sys
os.environ["UMB_GEMM_TRACE"] = "1"
os.environ["UMB_GEMM_TRACE_LAUNCH"] = "3"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("UMB_LIB_PATH", os.path.join(ROOT, "umbrella_b200", "libumbrella_b200.trace.so"))
import numpy as np, torch
sys.path.insert(0, ROOT)
from umbrella_b200 import ops, _lib
DEV = "cuda:0"
NAMES = ["entry", "setup", "w_issued", "dep_ok", "rec0_landed", "mma0", "mma_end", "acc0", "epi0_done", "exit",
         "prod_end", "epi_dep_ok", "acc1", "epi1_done", "red_acc_loaded", "red_reduced"]
M = int(os.environ.get("M", "13"))
for (K, N, mode, name) in ((8192, 10240, ops.UMB_EPI_STORE, "qkv"), (8192, 8192, ops.UMB_EPI_RESIDUAL, "o"),
                           (28672, 8192, ops.UMB_EPI_RESIDUAL, "down"), (8192, 57344, ops.UMB_EPI_SILU_MUL, "gate|up")):
    nbytes = ops.awq_packed_bytes(K, N)
    packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(8)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N // 2 if mode == ops.UMB_EPI_SILU_MUL else N, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    buf = np.zeros((1024, 16), dtype=np.uint64)
    for rep in range(2):                       # the first round warms up; the read resets the launch counter
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(8):
            ops.awq_gemm(x, packed[i], K, N, **kw)
        e.record()
        torch.cuda.synchronize()
        _lib.check(_lib.load().umb_debug_gemm_trace(buf.ctypes.data_as(C.c_void_p), 1024))
    t = buf[:148].astype(np.int64)
    live = t[:, 0] > 0
    dep = np.median(t[live, 3])
    print(f"== awq {name} M={M}: launch 3 of a chain of 8 ({s.elapsed_time(e) * 1000 / 8:.1f} us per launch), {live.sum()} CTAs; "
          f"us relative to dep_ok (median)")
    for i, n in enumerate(NAMES):
        col = t[:, i]
        ok = live & (col > 0)
        if ok.any():
            d = (col[ok] - dep) / 1e3
            print(f"   {n:14s} min {d.min():7.2f}  median {np.median(d):7.2f}  max {d.max():7.2f}  ({ok.sum()} CTAs)")
    del packed

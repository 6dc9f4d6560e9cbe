"""Per-CTA timeline of one GEMM launch (UMB_GEMM_TRACE=1): where the fixed per-launch time goes."""
import ctypes as C, os, sys
os.environ["UMB_GEMM_TRACE"] = "1"
os.environ.setdefault("UMB_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "umbrella_b200", "libumbrella_b200.trace.so"))  # python -m umbrella_b200.build --variant trace -DUMB_GEMM_TRACE_BUILD=1
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200 import ops, _lib
DEV = "cuda:0"
NAMES = ["entry", "setup", "w_issued", "dep_ok", "rec0_landed", "mma0", "mma_end", "acc0", "epi0_done", "exit",
         "prod_end", "epi_dep_ok", "acc1", "epi1_done", "red_acc_loaded", "red_reduced"]

def trace(label, launch, warm=3):
    for _ in range(warm):
        launch()
    torch.cuda.synchronize()
    launch()
    buf = np.zeros((148, 16), dtype=np.uint64)
    _lib.check(_lib.load().umb_debug_gemm_trace(buf.ctypes.data_as(C.c_void_p), 148))
    t = buf.astype(np.int64)
    live = t[:, 0] > t[:, 0].max() - 1_000_000          # CTAs of the most recent launch (grid may be < 148)
    t0 = t[live, 0].min()
    print(f"== {label}: {live.sum()} CTAs, kernel span {(t[live, 9].max() - t0) / 1e3:.2f} us (first entry -> last exit)")
    for i, n in enumerate(NAMES):
        col = t[:, i]
        ok = live & (col >= t0)
        if ok.any():
            d = (col[ok] - t0) / 1e3
            print(f"   {n:12s} min {d.min():7.2f}  median {np.median(d):7.2f}  max {d.max():7.2f} us  ({ok.sum()} CTAs)")

M = int(os.environ.get("M", "13"))
for (K, N, mode, name) in ((8192, 10240, ops.UMB_EPI_STORE, "qkv"), (8192, 8192, ops.UMB_EPI_RESIDUAL, "o"),
                           (8192, 57344, ops.UMB_EPI_SILU_MUL, "gate|up"), (28672, 8192, ops.UMB_EPI_RESIDUAL, "down")):
    nbytes = ops.awq_packed_bytes(K, N)
    packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(6)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N // 2 if mode == ops.UMB_EPI_SILU_MUL else N, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    it = [0]
    def launch():
        it[0] += 1
        ops.awq_gemm(x, packed[it[0] % 6], K, N, **kw)
    trace(f"awq {name} M={M} K={K} N={N}", launch)
    del packed
w = [(0.02 * torch.randn(3072, 2048, device=DEV)).to(torch.float16) for _ in range(40)]
x = torch.randn(3, 2048, device=DEV).to(torch.float16)
out = torch.zeros(3, 3072, dtype=torch.float16, device=DEV)
ws = ops.gemm_workspace(3, 2048, 3072, DEV)
it = [0]
def launch():
    it[0] += 1
    ops.gemm_f16(x, w[it[0] % 40], out=out, workspace=ws)
trace("fp16 draft qkv M=3 K=2048 N=3072", launch)

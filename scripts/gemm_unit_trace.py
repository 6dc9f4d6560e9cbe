"""Per-unit timeline of one CTA of an AWQ GEMM launch (trace build): which pipeline stage paces the steady state.
Stamps per CTA-local unit u: 0 weight TMA issued, 1 record seen landed by its dequant group, 2 dequant arithmetic of the
first chunk done (about to wait for the TMEM stage), 7 TMEM stage free, 3 A operand published (a_full), 4 MMA warp past
its waits, 5 MMAs issued + committed, 6 producer saw the unit retired (done)."""
import ctypes as C, os, sys
os.environ["UMB_GEMM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("UMB_LIB_PATH", os.path.join(ROOT, "umbrella_b200", "libumbrella_b200.trace.so"))
import numpy as np, torch
sys.path.insert(0, ROOT)
from umbrella_b200 import ops, _lib
DEV = "cuda:0"
M = int(os.environ.get("M", "13"))
for (K, N, mode, name) in ((8192, 57344, ops.UMB_EPI_SILU_MUL, "gate|up"), (8192, 8192, ops.UMB_EPI_RESIDUAL, "o")):
    nbytes = ops.awq_packed_bytes(K, N)
    packed = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(4)]
    x = torch.randn(M, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N // 2 if mode == ops.UMB_EPI_SILU_MUL else N, dtype=torch.float16, device=DEV)
    ws = ops.gemm_workspace(M, K, N, DEV)
    kw = dict(mode=mode, out=out, workspace=ws)
    if mode == ops.UMB_EPI_RESIDUAL:
        kw["residual"] = out
    for i in range(4):
        ops.awq_gemm(x, packed[i % 4], K, N, **kw)
    torch.cuda.synchronize()
    buf = np.zeros((1024, 16), dtype=np.uint64)
    _lib.check(_lib.load().umb_debug_gemm_trace(buf.ctypes.data_as(C.c_void_p), 1024))
    t = buf.reshape(-1)[8192:8192 + 256 * 8].reshape(256, 8).astype(np.int64)
    n_units = int((t[:, 5] > 0).sum())
    np.save(os.path.join(ROOT, "gpurun_out", f"unit_trace_{name.replace('|', '')}_M{M}.npy"), t[:n_units])
    t0 = t[0, 0]
    print(f"== {name} M={M}: CTA 0, {n_units} units; times in ns relative to the first TMA issue")
    print("unit   w_issue  landed  mma>a_full mma>full_x  a_full   mma_go  mma_done_issue  retired_seen | a_full-landed  mma_go-a_full  issue_dur  per-unit   (drain build: columns 3/4 are MMA-warp stamps)")
    prev = None
    for u in range(n_units):
        r = t[u]
        f = lambda v: f"{(v - t0):8d}" if v > 0 else "       -"
        per = (r[5] - prev) if prev is not None else 0
        prev = r[5]
        if u < 12 or u % 8 == 0 or u >= n_units - 6:
            print(f"{u:4d} {f(r[0])} {f(r[1])} {f(r[2])} {f(r[7])} {f(r[3])} {f(r[4])} {f(r[5])} {f(r[6])} | "
                  f"{r[3]-r[1]:8d} {r[4]-r[3]:8d} {r[5]-r[4]:8d} {per:8d}")
    mid = t[n_units // 4: 3 * n_units // 4]
    d = np.diff(mid[:, 5])
    print(f"steady state: {d.mean():.0f} ns per unit ({d.mean() * 1.965:.0f} cycles at 1965 MHz); "
          f"a_full - landed median {np.median(mid[:, 3] - mid[:, 1]):.0f} ns; mma_go - a_full median {np.median(mid[:, 4] - mid[:, 3]):.0f} ns; "
          f"MMA issue median {np.median(mid[:, 5] - mid[:, 4]):.0f} ns; landed - w_issue median {np.median(mid[:, 1] - mid[:, 0]):.0f} ns; "
          f"tmem wait median {np.median((mid[:, 7] - mid[:, 2])[mid[:, 7] > 0]):.0f} ns; retired_seen - mma_done median {np.median((mid[:, 6] - mid[:, 5])[mid[:, 6] > 0]):.0f} ns")
    del packed

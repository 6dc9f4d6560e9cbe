#!/usr/bin/env python
"""bench.py -- accepted tokens/sec at batch 1 of the speculative-decoding verify engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--no-reference-gpu] [--with-offload]

Workloads (BASELINE.json ``configs``) -- synthetic seeded weights of the real shapes, synthetic prompts,
scripted acceptance with the reference's DEFAULT_ACC (random weights make draft and target uncorrelated):

  static70b      (default; the configuration the metric is quoted on, configs[2] / greedy_config_48gb.json)
                 Llama-3.1-70B-AWQ-INT4 target + Llama-3.2-1B fp16 draft, sequoia_tree-3x4 (13 nodes), greedy verify,
                 prompt 512, generation 256, max_length 2048
  static8b       (configs[1] / code_config_24gb.json) Llama-3.1-8B fp16 target + 1B draft, 8b_sequoia_tree-5x6 (31 nodes)
  stochastic32k  (configs[3]) 70B-AWQ + 1B, T = 0.6 / top-p 0.9 / top-k 32 / repetition penalty 1.05 (chat_config_48gb),
                 prompt 30000, max_length 32768, KV pages of the upper layers resident in pinned host memory
  stochastic32k_hbm  the same with the whole KV cache in HBM
  static70b_b200tree the default workload on the packaged B200-tuned 16-node tree (sequoia_utils.generate_budget_tree)

A *step* is one speculative-decoding iteration (draft build_tree + target verify + accept + KV gather).  ``value`` =
accepted tokens / device time over exactly K steps with everything resident in HBM (CUDA events; max over ranks);
``e2e`` = the same metric through the public engine call ``engine.generate(input_ids=<host list>)`` (H2D of the prompt
from pinned memory, prefill, decode loop, D2H of the generated ids, wall clock) over 8 requests per replica.  N > 1:
one independent replica per GPU (no collective on the data path; torch.distributed only for the barrier and the
reduction of the timings).  At N = 1 the line also carries ``reference_gpu``: the reference's own GPU path
(baseline/ref_gpu.py: unmodified reference classes, vLLM's build of the upstream AWQ kernel, flashinfer) timed in the
same run on the same inputs, and ``cpu_baseline``: the oracle's C port on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T70 = "hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4"
T8 = "meta-llama/Llama-3.1-8B-Instruct"
D1 = "meta-llama/Llama-3.2-1B-Instruct"
WORKLOADS = {
    "static70b": dict(target=T70, draft=D1, tree="sequoia_tree-3x4.json", prompt=512, gen=256, max_length=2048,
                      sampling=dict(temperature=0.0),
                      text="Llama-3.1-70B-Instruct-AWQ-INT4 target + Llama-3.2-1B fp16 draft, sequoia_tree-3x4 (13 nodes), "
                           "greedy verify, prompt 512, generation 256, max_length 2048, one replica per GPU"),
    "static8b": dict(target=T8, draft=D1, tree="8b_sequoia_tree-5x6.json", prompt=512, gen=256, max_length=2048,
                     sampling=dict(temperature=0.0),
                     text="Llama-3.1-8B-Instruct fp16 target + Llama-3.2-1B fp16 draft, 8b_sequoia_tree-5x6 (31 nodes), "
                          "greedy verify, prompt 512, generation 256, max_length 2048, one replica per GPU"),
    "stochastic32k": dict(target=T70, draft=D1, tree="sequoia_tree-3x4.json", prompt=30000, gen=256, max_length=32768,
                          sampling=dict(temperature=0.6, topp=0.9, topk=32, repetition_penalty=1.05), kv_host_layers=40,
                          text="Llama-3.3-70B-Instruct-AWQ shape target + Llama-3.2-1B draft, sequoia_tree-3x4, stochastic "
                               "verify (T 0.6, top-p 0.9, top-k 32, repetition penalty 1.05), prompt 30000, generation 256, "
                               "max_length 32768, KV pages of layers 40..79 in pinned host memory"),
}
# the default workload on the packaged B200-tuned tree (sequoia_utils.tune_budget_tree(16, 10.0, 1.05): as many nodes as the
# 16-row padding of the verify pass holds, at the depth the measured costs favour) -- NOT the configuration the metric is quoted on
WORKLOADS["static70b_b200tree"] = dict(WORKLOADS["static70b"], tree="b200_sequoia_tree-16.json",
                                       text=WORKLOADS["static70b"]["text"].replace("sequoia_tree-3x4 (13 nodes)",
                                                                                   "b200_sequoia_tree-16 (16 nodes, depth 3)"))
# the same 32k-context stochastic workload with the whole KV cache resident in HBM (10.7 GB): what the paging costs
WORKLOADS["stochastic32k_hbm"] = dict(WORKLOADS["stochastic32k"], kv_host_layers=0,
                                      text=WORKLOADS["stochastic32k"]["text"].replace(
                                          "KV pages of layers 40..79 in pinned host memory", "KV cache resident in HBM"))
E2E_REQUESTS = 8          # per replica (2 for stochastic32k: a 30k-token prefill and PCIe-bound verifies per request)


def _json_file(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except Exception:
        return None


def _peaks():
    p = _json_file("MEASURED_PEAKS.json")
    if p and "hbm_gbs" in p:
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def build_engine(device: str, wl: dict):
    from umbrella_b200.models import llama as L
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    eng = StaticSpeculationEngine(wl["draft"], wl["target"], device=device, growmap_path=wl["tree"],
                                  max_length=wl["max_length"], **wl["sampling"])
    tshape, tawq = L.KNOWN_MODELS[wl["target"]]
    dshape, dawq = L.KNOWN_MODELS[wl["draft"]]
    kv_host = wl.get("kv_host_layers", 0)
    target = L.SyntheticLlama(tshape, max_length=wl["max_length"], awq=tawq, seed=1234, device=device, gen_device=device,
                              kv_host_layers=kv_host)
    draft = L.SyntheticLlama(dshape, max_length=wl["max_length"], awq=dawq, seed=4321, device=device, gen_device=device)
    eng.initialize(draft_model=draft, target_model=target)
    eng._set_eos([])                      # random logits must not end a request early
    return eng


def run_ours(args):
    import torch
    import torch.distributed as dist
    from umbrella_b200 import _lib
    from umbrella_b200.replicas import gather_over_replicas, reduce_over_replicas
    from umbrella_b200.sequoia_utils import expected_accept_length
    from umbrella_b200.speculation.scripted import ScriptedAcceptance

    wl = WORKLOADS[args.workload]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # the process group exists for the barrier and the reduction of the timings only (replicas, no data-path
        # collective); NCCL_DEBUG is left as the launcher set it so that its rank check can see the communicator
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    eng = build_engine(dev, wl)
    N = eng.tree_size
    script = ScriptedAcceptance(eng.Successors, seed=100 + rank, max_iters=4096, device=dev)
    eng.acceptance_script = script
    g = torch.Generator().manual_seed(rank)
    prompt = torch.randint(3, 30000, (1, wl["prompt"]), generator=g)          # examples/bench.py:31
    prompt_dev = prompt.to(dev)
    GEN = wl["gen"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def new_request():
        eng.reset()
        assert eng._prefill(prompt_dev)
        return eng.num_nodes

    def step():
        eng.build_tree()
        eng.verify()
        return eng.last_accept_length

    # ---- warm-up
    start = new_request()
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed: exactly K steps; a request is renewed (untimed prefill) after GEN tokens
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = _lib.launch_count()
    tokens, dev_ms = 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps_left = args.steps
    while steps_left > 0:
        if eng.num_nodes - start >= GEN or not eng.validate_status():
            start = new_request()
        e0.record()
        while steps_left > 0 and eng.num_nodes - start < GEN and eng.validate_status():
            tokens += step()
            steps_left -= 1
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    barrier()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel family, timed alone with CUDA events on the launching stream
    gemm_ms, gemm_bytes, gemm_launches, gemm_name = time_target_gemms(eng.target_model, N)
    verify_ms = time_verify(eng, prompt_dev)

    # ---- end to end through the public API, host buffers in and out
    eng.acceptance_script = script
    prompt_host = prompt[0].tolist()
    e2e_tokens, e2e_steps = 0, 0
    torch.cuda.synchronize()
    t0 = time.time()
    n_req = 2 if args.workload.startswith("stochastic32k") else E2E_REQUESTS
    for _ in range(n_req):
        out = eng.generate(input_ids=prompt_host, max_new_tokens=GEN)
        e2e_tokens += len(out["generated_tokens"])
        e2e_steps += out["large_model_steps"]
    torch.cuda.synchronize()
    e2e_s = time.time() - t0

    (dev_ms_max, e2e_s_max), (tokens_all, e2e_tokens_all) = reduce_over_replicas(
        [dev_ms, e2e_s], [float(tokens), float(e2e_tokens)], device=dev)
    per_replica = gather_over_replicas([tokens / (dev_ms * 1e-3), e2e_tokens / e2e_s], device=dev)

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = gemm_bytes / (gemm_ms * 1e-3) / 1e9
        wbytes = eng.target_model.weight_bytes()
        kv_row = 2 * eng.target_model.num_layers * eng.target_model.num_key_value_heads * eng.target_model.head_dim * 2
        kv_bytes = kv_row * (wl["prompt"] + GEN // 2 + 2 * N)
        traffic = _json_file("profiles", "r2_ncu_traffic.json") or _json_file("profiles", "r1_ncu_traffic.json") or {}
        line = {
            "metric": "accepted tokens/sec at bs=1 (Llama3.1-70B-AWQ target, static Sequoia tree, greedy verify)"
            if args.workload == "static70b" else f"accepted tokens/sec at bs=1 ({args.workload})",
            "value": tokens_all / (dev_ms_max * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (int4 weights, fp32 accumulate)" if eng.target_model.awq else "fp16 (fp32 accumulate)",
            "data": "synthetic (seeded random weights of the real shapes, random prompt, scripted acceptance DEFAULT_ACC)",
            "config": {"workload": wl["text"], "name": args.workload,
                       "l2": f"inputs larger than L2 ({wbytes / 1e9:.1f} GB of weights per verify pass)",
                       "mean_accept_len": tokens_all / (args.steps * world),
                       "expected_accept_len": expected_accept_length(eng.growmap)},
            "clocks": clocks,
            "e2e": {"value": e2e_tokens_all / e2e_s_max, "unit": "tokens/s",
                    "h2d_bytes_per_step": wl["prompt"] * 8 * n_req / max(e2e_steps, 1),
                    "d2h_bytes_per_step": (4 + N) * 4 + e2e_tokens * 8 / max(e2e_steps, 1),
                    "requests_per_replica": n_req,
                    "what": "engine.generate(input_ids=host list): H2D prompt, prefill, decode loop, D2H ids; wall clock"},
            "replicas": {"tokens_per_s_min": min(r[0] for r in per_replica), "tokens_per_s_max": max(r[0] for r in per_replica),
                         "e2e_tokens_per_s_min": min(r[1] for r in per_replica),
                         "e2e_tokens_per_s_max": max(r[1] for r in per_replica)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": gemm_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": traffic.get("awq_gemm_traffic_bytes_per_launch") if eng.target_model.awq else None,
                         "traffic_source": traffic.get("source", "profiles/r1_ncu_traffic.json: dram__bytes_read.sum + "
                                                       "dram__bytes_write.sum per launch, ncu --set full, mean of the 4 GEMM "
                                                       "shapes of a layer"),
                         "algorithmic_bytes_per_launch": gemm_bytes / gemm_launches,
                         "avg_launch_us": gemm_ms * 1e3 / gemm_launches},
            "verify_pass": {"ms": verify_ms, "algorithmic_bytes": wbytes + kv_bytes,
                            "hbm_frac": (wbytes + kv_bytes) / (verify_ms * 1e-3) / 1e9 / peak},
            "host": {"cpu_count": os.cpu_count()},
        }
        if getattr(eng.target_model, "kv_host_layers", 0):
            line["kv_paging"] = eng.target_model.kv_paging_stats()
        if world == 1:
            del eng
            torch.cuda.empty_cache()
            if args.with_offload and args.workload == "static70b":
                line["offload_num_cache_layers_0"] = time_offload(dev, wl, prompt_dev, tokens_all / args.steps)
            if not args.no_reference_gpu and args.workload in ("static70b", "static8b"):
                line["reference_gpu"] = reference_gpu(wl, min(args.steps, 20), line["value"], line["e2e"]["value"])
            if not args.no_cpu_baseline and args.workload == "static70b":
                line["cpu_baseline"] = cpu_baseline(tokens_all / args.steps)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def time_target_gemms(model, rows: int):
    """All weight-streaming GEMM launches of the decoder layers of one verify pass (4 per layer), nothing else."""
    import torch
    from umbrella_b200 import ops
    sh = model.config
    H, I = sh.hidden_size, sh.intermediate_size
    qkv_n = (model.num_heads + 2 * model.num_key_value_heads) * model.head_dim
    x = torch.randn(rows, H, device=model.device).to(torch.float16)
    xi = torch.randn(rows, I, device=model.device).to(torch.float16)
    y_qkv = torch.empty(rows, qkv_n, dtype=torch.float16, device=model.device)
    y_h = torch.zeros(rows, H, dtype=torch.float16, device=model.device)
    y_i = torch.empty(rows, I, dtype=torch.float16, device=model.device)
    ws = model._gemm_ws
    awq = model.awq

    def lin(xx, w, K, N, **kw):
        if awq:
            ops.awq_gemm(xx, w, K, N, workspace=ws, **kw)
        else:
            ops.gemm_f16(xx, w, workspace=ws, **kw)

    def one_pass():
        for lw in model.layer_weights:
            lin(x, lw["wqkv"], H, qkv_n, out=y_qkv)
            lin(x, lw["wo"], H, H, mode=ops.UMB_EPI_RESIDUAL, residual=y_h, out=y_h)
            lin(x, lw["wgu"], H, 2 * I, mode=ops.UMB_EPI_SILU_MUL, out=y_i)
            lin(xi, lw["wdown"], I, H, mode=ops.UMB_EPI_RESIDUAL, residual=y_h, out=y_h)

    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    s.record()
    for _ in range(reps):
        one_pass()
    e.record()
    torch.cuda.synchronize()
    nbytes = sum(lw[k].numel() * lw[k].element_size() for lw in model.layer_weights for k in ("wqkv", "wo", "wgu", "wdown"))
    n_launch = 4 * len(model.layer_weights)
    mp = 16 if rows <= 16 else (32 if rows <= 32 else 64)
    name = (f"gemm_kernel<{mp},true> (AWQ INT4 dequant + tcgen05 GEMM), {n_launch} launches per verify" if awq else
            f"gemm_kernel<{mp},false> (fp16 tcgen05 GEMM), {n_launch} launches per verify")
    return s.elapsed_time(e) / reps, nbytes, n_launch, name


def time_verify(eng, prompt_dev):
    """One target verify forward (the north-star pass) alone, CUDA events, mean of 10."""
    import torch
    eng.reset()
    eng._prefill(prompt_dev)
    n, N = eng.num_nodes, eng.tree_size
    ids, pos = eng.tokens[0, n:n + N], eng.position_ids[0, n:n + N]
    tm = eng.target_model
    times = []
    for i in range(13):
        tm.kv_cache.kv_offset = n
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        tm.verify_forward(ids, pos, n, eng.tree_bits, argmax_keys=eng._argmax_keys)
        e.record()
        torch.cuda.synchronize()
        eng._argmax_keys.zero_()
        if i >= 3:
            times.append(s.elapsed_time(e))
    eng.reset()
    return sum(times) / len(times)


def time_offload(dev, wl, prompt_dev, accept_len):
    """Our own paged target (LlamaAwqOffload, num_cache_layers = 0: every layer streamed from pinned host memory, the
    reference's variant A, llama.py:196-219): ms per verify and achieved H2D GB/s."""
    import torch
    from umbrella_b200.models import llama as L
    tshape, tawq = L.KNOWN_MODELS[wl["target"]]
    t0 = time.time()
    model = L.SyntheticLlama(tshape, max_length=wl["max_length"], awq=tawq, seed=1234, device=dev, gen_device=dev,
                             offload={"num_cache_layers": 0})
    setup = time.time() - t0
    P = prompt_dev.shape[1]
    model.prefill_forward(prompt_dev[0], torch.arange(P, device=dev), want_logits=False)
    N = 13
    ids = torch.randint(3, 30000, (N,), device=dev)
    pos = torch.arange(P, P + N, device=dev)
    bits = torch.tril(torch.ones(N, N, dtype=torch.int32, device=dev))
    bits = (bits << torch.arange(N, device=dev, dtype=torch.int32)[None]).sum(-1, keepdim=True).to(torch.int32)
    times = []
    for i in range(5):
        model.kv_cache.kv_offset = P
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        model.verify_forward(ids, pos, P, bits)
        e.record()
        torch.cuda.synchronize()
        if i >= 1:
            times.append(s.elapsed_time(e))
    ms = sum(times) / len(times)
    paged = model.paged_bytes_per_forward()
    out = {"ms_per_verify": ms, "h2d_gbs": paged / (ms * 1e-3) / 1e9, "paged_bytes_per_verify": paged,
           "tokens_per_s_upper_bound": accept_len / (ms * 1e-3), "setup_s": setup,
           "what": "umbrella_b200 LlamaAwqOffload, num_cache_layers=0: one pinned slab + one cudaMemcpyAsync per layer, "
                   "event-ordered double buffer; the draft's cost is not included"}
    del model
    torch.cuda.empty_cache()
    return out


def reference_gpu(wl, steps, ours_value, ours_e2e):
    """The reference's own GPU path in a subprocess (fresh interpreter: its `umbrella` package shadows ours)."""
    script = os.path.join(ROOT, "baseline", "ref_gpu.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "umbrella")):
        return {"unavailable": "baseline/_ref is absent (pip install --no-index --no-deps --target baseline/_ref <reference>)"}
    cmd = [sys.executable, script, "--steps", str(steps), "--warmup", "3", "--tree", wl["tree"], "--target", wl["target"],
           "--draft", wl["draft"], "--prompt-len", str(wl["prompt"]), "--gen-len", str(wl["gen"]),
           "--max-length", str(wl["max_length"])]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": f"baseline/ref_gpu.py rc={r.returncode}: {r.stderr.strip().splitlines()[-1:]}"}
        out = json.loads(lines[-1])
        out["ratio_ours_over_reference_gpu"] = ours_value / out["value"]
        out["e2e_ratio_ours_over_reference_gpu"] = ours_e2e / out["e2e_tokens_per_s"]
        return out
    except Exception as e:                                   # the reference arm must never take the product line down
        return {"unavailable": repr(e)[:300]}


def cpu_baseline(accept_len: float, passes: int = 3, threads: int | None = None) -> dict:
    from oracle.cpu_baseline import CpuLayerBaseline
    b = CpuLayerBaseline(threads=threads)
    b.layer_pass()                                  # warm-up (allocator, thread pool)
    secs = [b.layer_pass() for _ in range(passes)]
    layer_s = sum(secs) / len(secs)
    return {"value": b.tokens_per_second(layer_s, accept_len), "unit": "tokens/s", "cores": b.threads, "kind": "port",
            "sample": f"{passes} timed passes of ONE 70B-AWQ decoder layer (7 INT4 linears through the C oracle oracle/awq_cpu.c "
                      f"-- weights stay INT4, one pass, no dequantised matrix -- norms, RoPE, attention over 512+13 rows, 1/80 of "
                      f"the lm_head) on {b.threads} host threads, {layer_s:.3f} s per layer; tokens/s = accept_len / (80 x layer "
                      f"time), draft cost ignored",
            "layer_seconds": layer_s}


def run_reference(args):
    """Reference arm (driver contract): the path's arithmetic on the box's host cores.  The reference has no CPU
    implementation of its own (it is GPU compute with optional host-RAM paging) and no compilable sources, so this is
    the oracle's C port (oracle/awq_cpu.c + torch CPU glue), all host threads.  A step = ONE decoder layer of the
    verify forward (1/80 of the workload's step, stated in `sample`); `value` extrapolates to the whole verify."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuLayerBaseline
    from umbrella_b200.sequoia_utils import expected_accept_length, load_growmap
    wl = WORKLOADS["static70b"]
    b = CpuLayerBaseline(threads=os.cpu_count())
    for _ in range(max(args.warmup, 1)):
        b.layer_pass()
    secs = [b.layer_pass() for _ in range(args.steps)]
    layer_s = sum(secs) / len(secs)
    acc = expected_accept_length(load_growmap(wl["tree"]))
    v = b.tokens_per_second(layer_s, acc)
    sample = (f"each step = ONE 70B-AWQ decoder layer of the verify forward (1/80 of a verify) on {b.threads} host threads "
              f"({layer_s:.3f} s/layer); value = accept length {acc:.3f} (scripted expectation) / (80 x layer time)")
    print(json.dumps({
        "impl": "reference", "metric": "accepted tokens/sec at bs=1 (Llama3.1-70B-AWQ target, static Sequoia tree, greedy verify)",
        "value": v, "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": layer_s * 1e3, "ms_per_verify_extrapolated": layer_s * 80 * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (int4 weights, fp32 accumulate)",
        "data": "synthetic (seeded random weights of the real shapes)",
        "config": {"workload": wl["text"], "name": "static70b", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": b.threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="static70b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--with-offload", action="store_true", help="also time our paged 70B target (num_cache_layers=0)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when invoked bare with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path: no GPU is visible (there is no CPU fallback)")
    run_ours(args)


if __name__ == "__main__":
    main()

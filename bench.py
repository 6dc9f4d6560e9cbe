#!/usr/bin/env python
"""bench.py -- accepted tokens/sec at batch 1 of the speculative-decoding verify engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric): Llama-3.1-70B-AWQ-INT4 target + Llama-3.2-1B fp16
draft, Sequoia static tree ``sequoia_tree-3x4`` (13 nodes), greedy verify, synthetic seeded weights
of the real shapes, synthetic prompt of 512 tokens, generation_length 256, scripted acceptance
with the reference's DEFAULT_ACC (random weights make draft and target uncorrelated).

A *step* is one speculative-decoding iteration (draft build_tree + target verify + accept + KV
gather).  ``value`` = accepted tokens / device time over exactly K steps with everything resident in
HBM (CUDA events; max over ranks); ``e2e`` = the same metric through the public engine call
``engine.generate(input_ids=<host list>)`` (H2D of the prompt from pinned memory, prefill, decode
loop, D2H of the generated ids, wall clock).  N > 1: one independent replica per GPU (no collective on
the data path; torch.distributed is used for the barrier and the reduction of the timings only).
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

TREE = "sequoia_tree-3x4.json"
PROMPT_LEN = 512
GEN_LEN = 256
MAX_LENGTH = 2048


def _ncu_traffic():
    """DRAM bytes per AWQ GEMM launch from the committed ncu --set full capture (scripts/ncu_traffic.py);
    None when the capture is absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        return float(t["awq_gemm_traffic_bytes_per_launch"])
    except Exception:
        return None


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
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


def build_engine(device: str, seed: int):
    import torch
    from umbrella_b200.models import llama as L
    from umbrella_b200.speculation.static_speculation_engine import StaticSpeculationEngine
    eng = StaticSpeculationEngine("meta-llama/Llama-3.2-1B-Instruct", "hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4",
                                  device=device, growmap_path=TREE, max_length=MAX_LENGTH, temperature=0.0)
    target = L.SyntheticLlama(L.LLAMA31_70B, max_length=MAX_LENGTH, awq=True, seed=1234, device=device, gen_device=device)
    draft = L.SyntheticLlama(L.LLAMA32_1B, max_length=MAX_LENGTH, awq=False, seed=4321, device=device, gen_device=device)
    eng.initialize(draft_model=draft, target_model=target)
    eng._set_eos([])                      # random logits must not end a request early
    return eng


def run_ours(args):
    import torch
    import torch.distributed as dist
    from umbrella_b200 import _lib
    from umbrella_b200.sequoia_utils import expected_accept_length
    from umbrella_b200.speculation.scripted import ScriptedAcceptance

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ["NCCL_DEBUG"] = "NONE"       # NCCL prints its version banner to stdout at VERSION and WARN level; rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    eng = build_engine(dev, seed=rank)
    N = eng.tree_size
    script = ScriptedAcceptance(eng.Successors, seed=100 + rank, max_iters=4096, device=dev)
    eng.acceptance_script = script
    g = torch.Generator().manual_seed(rank)
    prompt = torch.randint(3, 30000, (1, PROMPT_LEN), generator=g)          # examples/bench.py:31
    prompt_dev = prompt.to(dev)

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
    # ---- timed: exactly K steps; a request is renewed (untimed prefill) after GEN_LEN tokens
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = _lib.launch_count()
    tokens, dev_ms = 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps_left = args.steps
    wall0 = time.time()
    while steps_left > 0:
        if eng.num_nodes - start >= GEN_LEN or not eng.validate_status():
            start = new_request()
        e0.record()
        while steps_left > 0 and eng.num_nodes - start < GEN_LEN and eng.validate_status():
            tokens += step()
            steps_left -= 1
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    barrier()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel family (AWQ GEMM): its 4 launches per layer, all 80 layers,
    #      timed alone with CUDA events on the launching stream (weights 35.6 GB >> L2)
    awq_ms, awq_bytes = time_awq_gemms(eng.target_model, N)
    verify_ms = time_verify(eng, prompt_dev)

    # ---- end to end through the public API, host buffers in and out
    eng.acceptance_script = script
    prompt_host = prompt[0].tolist()
    t0 = time.time()
    e2e_reqs = 2
    e2e_tokens, e2e_steps = 0, 0
    for _ in range(e2e_reqs):
        out = eng.generate(input_ids=prompt_host, max_new_tokens=GEN_LEN)
        e2e_tokens += len(out["generated_tokens"])
        e2e_steps += out["large_model_steps"]
    torch.cuda.synchronize()
    e2e_s = time.time() - t0

    from umbrella_b200.replicas import reduce_over_replicas
    (dev_ms_max, e2e_s_max), (tokens_all, e2e_tokens_all) = reduce_over_replicas(
        [dev_ms, e2e_s], [float(tokens), float(e2e_tokens)], device=dev)

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = awq_bytes / (awq_ms * 1e-3) / 1e9
        wbytes = eng.target_model.weight_bytes()
        kv_bytes = 327680 * (PROMPT_LEN + GEN_LEN // 2 + 2 * N)
        line = {
            "metric": "accepted tokens/sec at bs=1 (Llama3.1-70B-AWQ target, static Sequoia tree, greedy verify)",
            "value": tokens_all / (dev_ms_max * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (int4 weights, fp32 accumulate)",
            "data": "synthetic (seeded random weights of the real shapes, random prompt, scripted acceptance DEFAULT_ACC)",
            "config": {"workload": "Llama-3.1-70B-Instruct-AWQ-INT4 target + Llama-3.2-1B fp16 draft, sequoia_tree-3x4 "
                                   "(13 nodes), prompt 512, generation 256, max_length 2048, one replica per GPU",
                       "l2": "inputs larger than L2 (37.7 GB of weights per verify pass)",
                       "mean_accept_len": tokens_all / (args.steps * world),
                       "expected_accept_len": expected_accept_length(eng.growmap)},
            "clocks": clocks,
            "e2e": {"value": e2e_tokens_all / e2e_s_max, "unit": "tokens/s",
                    "h2d_bytes_per_step": PROMPT_LEN * 8 * e2e_reqs / max(e2e_steps, 1),
                    "d2h_bytes_per_step": (4 + N) * 4 + e2e_tokens * 8 / max(e2e_steps, 1),
                    "what": "engine.generate(input_ids=host list): H2D prompt, prefill, decode loop, D2H ids; wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "gemm_kernel<16,true> (AWQ INT4 dequant + tcgen05 GEMM), 320 launches per verify",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": _ncu_traffic(),
                         "traffic_source": "profiles/r1_ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                           "ncu --set full, mean of the 4 GEMM shapes of a layer",
                         "algorithmic_bytes_per_launch": awq_bytes / 320, "avg_launch_us": awq_ms * 1e3 / 320},
            "verify_pass": {"ms": verify_ms, "algorithmic_bytes": wbytes + kv_bytes,
                            "hbm_frac": (wbytes + kv_bytes) / (verify_ms * 1e-3) / 1e9 / peak},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(tokens_all / args.steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_awq_gemms(model, rows: int):
    """All AWQ GEMM launches of one verify pass (4 per layer x L layers), nothing else, CUDA events."""
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

    def one_pass():
        for lw in model.layer_weights:
            ops.awq_gemm(x, lw["wqkv"], H, qkv_n, out=y_qkv, workspace=ws)
            ops.awq_gemm(x, lw["wo"], H, H, mode=ops.UMB_EPI_RESIDUAL, residual=y_h, out=y_h, workspace=ws)
            ops.awq_gemm(x, lw["wgu"], H, 2 * I, mode=ops.UMB_EPI_SILU_MUL, out=y_i, workspace=ws)
            ops.awq_gemm(xi, lw["wdown"], I, H, mode=ops.UMB_EPI_RESIDUAL, residual=y_h, out=y_h, workspace=ws)

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
    nbytes = sum(lw[k].numel() for lw in model.layer_weights for k in ("wqkv", "wo", "wgu", "wdown"))
    return s.elapsed_time(e) / reps, nbytes


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


def cpu_baseline(accept_len: float, passes: int = 2, threads: int | None = None) -> dict:
    from oracle.cpu_baseline import CpuLayerBaseline
    b = CpuLayerBaseline(threads=threads)
    b.layer_pass()                                  # warm-up (allocator, thread pool)
    secs = [b.layer_pass() for _ in range(passes)]
    layer_s = sum(secs) / len(secs)
    return {"value": b.tokens_per_second(layer_s, accept_len), "unit": "tokens/s", "cores": b.threads, "kind": "port",
            "sample": f"{passes} timed passes of ONE 70B-AWQ decoder layer (7 INT4 linears dequantised + multiplied, norms, "
                      f"RoPE, attention over 512+13 rows, 1/80 of the lm_head) on {b.threads} host threads, "
                      f"{layer_s:.2f} s per layer; tokens/s = accept_len / (80 x layer time), draft cost ignored",
            "layer_seconds": layer_s}


def run_reference(args):
    """Reference arm: the path's arithmetic on the box's host cores (oracle port; the reference has no
    CPU implementation of its own and none of its GPU dependencies exist for this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuLayerBaseline
    import torch
    b = CpuLayerBaseline(threads=os.cpu_count())
    for _ in range(min(args.warmup, 1)):
        b.layer_pass()
    steps = max(1, min(args.steps, 3))
    t0 = time.time()
    secs = [b.layer_pass() for _ in range(steps)]
    layer_s = sum(secs) / len(secs)
    from umbrella_b200.sequoia_utils import expected_accept_length, load_growmap
    acc = expected_accept_length(load_growmap(TREE))
    v = b.tokens_per_second(layer_s, acc)
    sample = (f"{steps} passes of ONE 70B-AWQ decoder layer of the verify forward on {b.threads} host threads "
              f"({layer_s:.2f} s/layer), extrapolated x80 layers; accept length = scripted expectation {acc:.3f}")
    print(json.dumps({
        "impl": "reference", "metric": "accepted tokens/sec at bs=1 (Llama3.1-70B-AWQ target, static Sequoia tree, greedy verify)",
        "value": v, "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": layer_s * 80 * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 (int4 weights, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": "Llama-3.1-70B-Instruct-AWQ-INT4 verify, sequoia_tree-3x4 (13 nodes), prefix 512; CPU port"},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": b.threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

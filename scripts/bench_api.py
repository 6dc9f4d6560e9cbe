"""What the request API costs on top of the engine: the default bench workload (70B-AWQ target + 1B
draft, 3x4 tree, scripted acceptance, prompt 512 / generation 256) served over the socket -- APIClient -> ReplicaFront ->
APIServer (engine on the server's worker thread) -- against the same requests through a direct engine.generate().

    python scripts/bench_api.py [--requests 8] [--workload static70b]      # one JSON line
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from umbrella_b200.api.client import APIClient  # noqa: E402
from umbrella_b200.api.front import ReplicaFront  # noqa: E402
from umbrella_b200.api.server import APIServer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, default=8)
ap.add_argument("--workload", default="static70b", choices=sorted(bench.WORKLOADS))
args = ap.parse_args()
wl = bench.WORKLOADS[args.workload]


def factory(device, config):
    from umbrella_b200.speculation.scripted import ScriptedAcceptance
    eng = bench.build_engine(device, wl)
    eng.acceptance_script = ScriptedAcceptance(eng.Successors, seed=100, max_iters=4096, device=device)
    return eng


import torch  # noqa: E402

t0 = time.time()
server = APIServer({}, device="cuda:0", port=0, engine_factory=factory).start()
front = ReplicaFront([("127.0.0.1", server.port)], port=0).start()
setup_s = time.time() - t0
g = torch.Generator().manual_seed(0)
prompt = torch.randint(3, 30000, (wl["prompt"],), generator=g).tolist()
request = dict(input_ids=prompt, max_new_tokens=wl["gen"], **wl["sampling"])


def run(call):
    call(**request)                                    # warm-up: graph capture of every bucket this request touches
    tokens, steps = 0, 0
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(args.requests):
        out = call(**request)
        tokens += len(out["generated_tokens"])
        steps += out["large_model_steps"]
    torch.cuda.synchronize()
    s = time.time() - t
    return {"tokens_per_s": tokens / s, "seconds": s, "tokens": tokens, "steps": steps}


with APIClient(port=front.port, retry_seconds=0.2, max_retries=100) as client:
    over_socket = run(client.get_output)
with APIClient(port=server.port, retry_seconds=0.2, max_retries=100) as client:
    one_hop = run(client.get_output)
direct = run(server.engine.generate)
payload = len(__import__("pickle").dumps(request))
line = {"what": "request API overhead, " + wl["text"], "requests": args.requests, "setup_s": setup_s,
        "request_pickle_bytes": payload,
        "client_front_server": over_socket, "client_server": one_hop, "direct_generate": direct,
        "api_over_direct": over_socket["tokens_per_s"] / direct["tokens_per_s"],
        "ms_per_request_added_by_the_api": 1e3 * (over_socket["seconds"] - direct["seconds"]) / args.requests}
print(json.dumps(line), flush=True)
front.shutdown()
server.shutdown()

"""The request API over several GPUs (the reference's app/api.py runs one engine behind one port; SURVEY.md 8e / 8f-4:
one replica per GPU behind one front, no data-path collective).

    python examples/api_replicas.py --server --gpus 8 --port 65432 [--config configs/greedy_config_48gb.json]
    python examples/api_replicas.py --port 65432 --requests 16 --clients 8          # any umbrella.api client works

The server side starts one ``APIServer`` process per GPU (ports ``port+1 ..``), waits until every replica has built its
engine, then serves the reference's wire protocol on ``port`` and places each request on the least-loaded replica.
Without checkpoints the models named in the config carry seeded synthetic weights and requests are token-id lists.
"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella.api.client import APIClient  # noqa: E402  (alias of umbrella_b200)
from umbrella.api.front import ReplicaFront, launch_replicas  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--server", action="store_true")
ap.add_argument("--config", default=None)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--port", type=int, default=65432)
ap.add_argument("--real-weights", action="store_true")
ap.add_argument("--requests", type=int, default=8)
ap.add_argument("--clients", type=int, default=4)
ap.add_argument("--prompt-len", type=int, default=256)
ap.add_argument("--max-new-tokens", type=int, default=128)
args = ap.parse_args()

if args.server:
    config = {"model": "hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4",
              "draft_model": "meta-llama/Llama-3.2-1B-Instruct", "engine": "static",
              "growmap_path": "sequoia_tree-3x4.json", "max_length": 2048, "temperature": 0.0}
    if args.config:
        config = json.load(open(args.config))
        if "growmap_path" in config:
            config["growmap_path"] = os.path.basename(config["growmap_path"])
    if not args.real_weights:
        config["synthetic"] = True
    replicas = launch_replicas(config, devices=[f"cuda:{i}" for i in range(args.gpus)], base_port=args.port + 1)
    try:
        ReplicaFront(replicas.addresses(), port=args.port).run()
    finally:
        replicas.stop()
else:
    import random
    done, lock = [], threading.Lock()

    def worker(k: int) -> None:
        rng = random.Random(k)
        with APIClient(port=args.port) as client:
            for _ in range(k, args.requests, args.clients):
                ids = [rng.randrange(3, 30000) for _ in range(args.prompt_len)]
                out = client.get_output(input_ids=ids, max_new_tokens=args.max_new_tokens, temperature=0.0)
                with lock:
                    done.append(out)

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(args.clients)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    wall = time.perf_counter() - t0
    ok = [o for o in done if o.get("processed")]
    tokens = sum(len(o.get("generated_tokens", [])) for o in ok)
    print(json.dumps({"requests": len(done), "processed": len(ok), "tokens": tokens, "wall_s": round(wall, 3),
                      "tokens_per_s": round(tokens / wall, 1),
                      "per_replica": {str(r): sum(1 for o in ok if o.get("replica") == r)
                                      for r in sorted({o.get("replica") for o in ok})}}))

"""One-GPU check of the replica path end to end with the product's own engine factory: launch_replicas spawns an
APIServer process that builds AutoEngine.from_config (synthetic weights) on its worker thread, the front serves two
concurrent clients over the reference wire protocol.      GPUS=1 python scripts/check_api_replicas.py"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella_b200.api.client import APIClient  # noqa: E402
from umbrella_b200.api.front import ReplicaFront, launch_replicas  # noqa: E402

if __name__ == "__main__":
    gpus = int(os.environ.get("GPUS", "1"))
    config = {"model": os.environ.get("TARGET", "hugging-quants/Meta-Llama-3.1-8B-Instruct-AWQ-INT4"),
              "draft_model": "meta-llama/Llama-3.2-1B-Instruct", "engine": "static",
              "growmap_path": "sequoia_tree-3x4.json", "max_length": 2048, "temperature": 0.0, "synthetic": True}
    t0 = time.time()
    replicas = launch_replicas(config, devices=[f"cuda:{i}" for i in range(gpus)], startup_timeout_s=300)
    front = ReplicaFront(replicas.addresses(), port=0).start()
    setup_s = time.time() - t0
    outs, lock = [], threading.Lock()

    def client(k):
        import random
        rng = random.Random(k)
        with APIClient(port=front.port, retry_seconds=0.2, max_retries=50) as c:
            for _ in range(2):
                out = c.get_output(input_ids=[rng.randrange(3, 30000) for _ in range(256)], max_new_tokens=64,
                                   temperature=0.0)
                with lock:
                    outs.append(out)

    t1 = time.time()
    threads = [threading.Thread(target=client, args=(k,)) for k in range(2 * gpus)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    wall = time.time() - t1
    ok = [o for o in outs if o.get("processed")]
    print(json.dumps({"replicas": gpus, "setup_s": round(setup_s, 1), "requests": len(outs), "processed": len(ok),
                      "tokens": sum(len(o["generated_tokens"]) for o in ok), "wall_s": round(wall, 3),
                      "per_replica": front.stats()["served"],
                      "errors": [o.get("response") for o in outs if not o.get("processed")][:2],
                      "avg_accept_tokens": [round(o["avg_accept_tokens"], 2) for o in ok]}), flush=True)
    front.shutdown()
    replicas.stop()

"""Speculative decoding through the reference's front-end surface (mirror of the reference's
examples/spec_generate.py): ``AutoEngine.from_config(device, **config)`` -> ``initialize()`` -> ``generate(...)``.

    python examples/spec_generate.py --config configs/greedy_config_48gb.json [--max-new-tokens 128]

There are no checkpoints in this environment, so the models named in the config are instantiated with seeded
synthetic weights of the real architectures (umbrella_b200/models/auto_model.py) and the prompt is a list of
random token ids; with real weights the same calls take ``context="..."`` text.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella.speculation.auto_engine import AutoEngine  # noqa: E402  (alias of umbrella_b200)

ap = argparse.ArgumentParser()
ap.add_argument("--config", default=None, help="a reference configs/*.json; default: the 70B-AWQ + 1B static config")
ap.add_argument("--max-new-tokens", type=int, default=128)
ap.add_argument("--prompt-len", type=int, default=256)
ap.add_argument("--real-weights", action="store_true", help="load the safetensors checkpoints (local dir / HF cache) "
                "instead of seeded synthetic weights of the same architectures")
args = ap.parse_args()

config = {"model": "hugging-quants/Meta-Llama-3.1-70B-Instruct-AWQ-INT4", "draft_model": "meta-llama/Llama-3.2-1B-Instruct",
          "engine": "static", "growmap_path": "sequoia_tree-3x4.json", "max_length": 2048, "temperature": 0.0}
if args.config:
    config = json.load(open(args.config))
    if "growmap_path" in config:
        config["growmap_path"] = os.path.basename(config["growmap_path"])      # shipped trees are looked up by name

if not args.real_weights:
    config["synthetic"] = True                                                # opt-in: no checkpoints offline
engine = AutoEngine.from_config(device="cuda:0", **config)
engine.initialize()
engine._set_eos([])                                                           # random weights: do not stop on a stray EOS
prompt = torch.randint(3, 30000, (args.prompt_len,)).tolist()
out = engine.generate(input_ids=prompt, max_new_tokens=args.max_new_tokens, temperature=0.0)
print(json.dumps({k: out[k] for k in ("avg_accept_tokens", "time_per_output_token", "large_model_steps")}))
print("first tokens:", out["generated_tokens"][:16])

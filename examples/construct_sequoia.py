"""Measure a draft / target pair's acceptance-rate vector and build Sequoia trees from it (the reference's
examples/construct_sequoia.py over ``umbrella.sequoia_utils``: there the measuring loop lives in the script and
``measure_acceptance_rate`` is empty; here the script only parses arguments).

    python examples/construct_sequoia.py --w 3 --d 4 --dst my_tree.json [--data pairs.jsonl] [--temperature 0.6]
    python examples/construct_sequoia.py --budget 16 --fixed-ms 10.0 --draft-forward-ms 1.05     # B200 cost-model tuner

``--data``: JSON lines with ``prompt_ids`` / ``solution_ids`` token lists (or ``prompt`` / ``canonical_solution`` text when
the tokenizer of ``--model`` is available locally, as for HumanEval in the reference).  Without ``--data`` the measurement
runs on random token pairs, which with synthetic weights only exercises the plumbing: uncorrelated random models accept
nothing, so the tree is then built from DEFAULT_ACC.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umbrella.models.auto_model import AutoModelLM  # noqa: E402  (alias of umbrella_b200)
from umbrella.sequoia_utils import (DEFAULT_ACC, expected_accept_length, generate_sequoia_tree,  # noqa: E402
                                    measure_acceptance_rate, tune_budget_tree)
from umbrella.templates import Prompts, SysPrompts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="meta-llama/Llama-3.1-8B-Instruct")
ap.add_argument("--draft_model", default="meta-llama/Llama-3.2-1B-Instruct")
ap.add_argument("--w", type=int, default=3, help="tree width")
ap.add_argument("--d", type=int, default=4, help="tree depth")
ap.add_argument("--dst", default="sequoia_tree.json")
ap.add_argument("--data", default=None)
ap.add_argument("--samples", type=int, default=8)
ap.add_argument("--temperature", type=float, default=0.0)
ap.add_argument("--topp", type=float, default=0.9)
ap.add_argument("--repetition_penalty", type=float, default=1.0)
ap.add_argument("--real-weights", action="store_true")
ap.add_argument("--budget", type=int, default=0, help="> 0: also tune a budget tree of this many nodes (B200 cost model)")
ap.add_argument("--fixed-ms", type=float, default=10.0, help="verify + accept + gathers per iteration")
ap.add_argument("--draft-forward-ms", type=float, default=1.05)
args = ap.parse_args()

DEVICE, MAX_LEN = "cuda:0", 2048
models = []
for name in (args.draft_model, args.model):
    m = AutoModelLM.from_pretrained(model_name=name, offload=False, cuda_graph=False, batch_size=1, max_length=MAX_LEN,
                                    dtype=torch.float16, device=DEVICE)
    m.alloc(**({} if args.real_weights else {"synthetic": True}))
    models.append(m)
draft_model, target_model = models

pairs = []
if args.data:
    tokenizer = None
    for line in open(args.data):
        rec = json.loads(line)
        if "prompt_ids" in rec:
            pairs.append((rec["prompt_ids"], rec["solution_ids"]))
            continue
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(args.model, local_files_only=True)
        context = SysPrompts["llama3-code"] + Prompts["llama3-code"].format("Help me complete the code.") + rec["prompt"]
        pairs.append((tokenizer.encode(context), tokenizer.encode(rec["canonical_solution"])[1:]))      # BOS dropped
else:
    g = torch.Generator().manual_seed(0)
    pairs = [(torch.randint(3, 30000, (96,), generator=g), torch.randint(3, 30000, (64,), generator=g))
             for _ in range(args.samples)]

width = max(args.w, 4)
acc = measure_acceptance_rate(draft_model, target_model, pairs, temperature=args.temperature, topp=args.topp,
                              repetition_penalty=args.repetition_penalty, width=width, max_length=MAX_LEN)
print("acceptance rate by draft rank:", [round(a, 4) for a in acc])
if sum(acc) < 1e-3:
    print("nothing accepted (uncorrelated models): building from DEFAULT_ACC", DEFAULT_ACC)
    acc = list(DEFAULT_ACC)
acc = [max(a, 1e-6) for a in acc]
gm = generate_sequoia_tree(width=args.w, depth=args.d, acc=acc, json_file=args.dst)
print(f"{args.dst}: {gm['size']} nodes, expected tokens per verify {expected_accept_length(gm, acc):.3f}")
if args.budget > 0:
    best, rate = tune_budget_tree(args.budget, args.fixed_ms, args.draft_forward_ms, acc=acc)
    dst = os.path.splitext(args.dst)[0] + f"-budget{args.budget}.json"
    with open(dst, "w") as f:
        json.dump(best, f)
    print(f"{dst}: {best['size']} nodes, depth {max(best['depth'])}, expected tokens per verify "
          f"{expected_accept_length(best, acc):.3f}, predicted {rate:.1f} tokens/s")

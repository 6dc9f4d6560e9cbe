"""TEST INFRASTRUCTURE.  tests/golden/stream.json: what the reference's own ``generate_stream``
(umbrella/speculation/static_speculation_engine.py:437-565, dynamic...:412-540) yields, run on CPU over an engine object
whose model-facing methods are scripted (``verify`` commits a scripted number of tokens per step; the tokenizer maps a
token id to a fixed word piece).  The product's ``speculation_utils.stream_generation`` must release the same texts, and its ``speculative_decoding``
loops must print and return the same (static...:211-255).
      python oracle/make_golden_stream.py
"""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import OUT, import_reference  # noqa: E402

VOCAB = ["ab", " cd", "e", " f", "g.", " hi", "j", " k!", " lm", "n", " o", "p?"]
START = 10
CASES = [   # per_step, max_new_tokens, verify() returns False at this step (None: never), validate_status room, stop_distance
    dict(per_step=[3, 1, 4, 2], max_new=40, stop_after=None, room=10_000, stop_distance=4),
    dict(per_step=[2], max_new=40, stop_after=5, room=10_000, stop_distance=4),
    dict(per_step=[5, 3], max_new=64, stop_after=None, room=10_000, stop_distance=4),
    dict(per_step=[1], max_new=2, stop_after=None, room=10_000, stop_distance=4),
    dict(per_step=[4], max_new=200, stop_after=None, room=START + 30, stop_distance=4),
    dict(per_step=[2, 2, 3], max_new=48, stop_after=None, room=10_000, stop_distance=16),
    dict(per_step=[13], max_new=100, stop_after=3, room=10_000, stop_distance=32),
]


class Tok:
    def decode(self, ids, **kwargs):
        return "".join(VOCAB[i] for i in ids)


def scripted(cls, case):
    eng = object.__new__(cls)
    eng.device = "cpu"
    eng.stop_distance = case["stop_distance"]
    eng.tokens = torch.zeros(1, 4096, dtype=torch.int64)
    eng.num_nodes = 0
    eng.tokenizer = Tok()
    eng.step = 0

    def prefill(input_ids):
        eng.num_nodes = START
        return True

    def verify():
        k = case["per_step"][eng.step % len(case["per_step"])]
        for i in range(k):
            eng.tokens[0, eng.num_nodes + i] = (eng.num_nodes + i) % len(VOCAB)
        eng.num_nodes += k
        eng.step += 1
        return case["stop_after"] is None or eng.step < case["stop_after"]

    eng._prefill = prefill
    eng.build_tree = lambda: None
    eng.verify = verify
    eng.validate_status = lambda: eng.num_nodes <= case["room"]
    eng.update_generation_args = lambda **kw: None
    eng.final = None

    def reset():
        eng.final = eng.num_nodes
    eng.reset = reset
    return eng


def main():
    import_reference()
    import umbrella.speculation.dynamic_speculation_engine as dse
    import umbrella.speculation.static_speculation_engine as sse
    torch.cuda.synchronize = lambda *a, **k: None            # the loop brackets its timers with it; CPU here
    out = []
    for case in CASES:
        per_engine = []
        for cls in (sse.StaticSpeculationEngine, dse.DynamicSpeculationEngine):
            eng = scripted(cls, case)
            chunks = list(eng.generate_stream(input_ids=[1, 2, 3], max_new_tokens=case["max_new"]))
            per_engine.append(([t for t, _ in chunks], eng.final - START))
        assert per_engine[0] == per_engine[1], "the two reference engines stream differently"
        texts, committed = per_engine[0]
        # the interactive loop (static...:211-255, dynamic...:167-213): what it prints and returns
        printed = []
        for cls in (sse.StaticSpeculationEngine, dse.DynamicSpeculationEngine):
            eng = scripted(cls, case)
            eng._prefill(None)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                dec_len, _, steps = eng.speculative_decoding(max_new_tokens=case["max_new"])
            printed.append((buf.getvalue(), dec_len, steps))
        assert printed[0] == printed[1], "the two reference engines print differently"
        # one API request (static...:373-434, dynamic...:343-409): the result keys the server sends back
        results = []
        for cls in (sse.StaticSpeculationEngine, dse.DynamicSpeculationEngine):
            eng = scripted(cls, case)
            r = eng.generate(input_ids=[1, 2, 3], max_new_tokens=case["max_new"], temperature=0.0)
            r.pop("time_per_output_token")
            results.append(r)
        assert results[0] == results[1], "the two reference engines answer differently"
        out.append(dict(case, texts=texts, committed=committed, printed=printed[0][0], dec_len=printed[0][1],
                        steps=printed[0][2], generate=results[0]))
        print(case, "->", len(texts), "chunks,", committed, "tokens")
    with open(os.path.join(OUT, "stream.json"), "w") as f:
        json.dump({"vocab": VOCAB, "start": START, "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()

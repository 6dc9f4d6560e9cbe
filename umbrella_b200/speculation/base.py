"""BaseEngine: the engine contract the front-ends program against.

The reference states it as thirteen abstract methods (umbrella/speculation/base.py:4-60).  Here the contract is
a table -- method name -> what a front-end may rely on -- enforced when an engine class is *defined*: a subclass
that lacks one of the methods cannot be created, which is a stricter form of the reference's "cannot be
instantiated".
"""
from __future__ import annotations

ENGINE_CONTRACT = {
    # set-up and state
    "initialize": "allocate buffers, build both models, load the tokenizer",
    "reset": "zero the token / position state and clear both KV caches (buffers are kept)",
    "validate_status": "False once num_nodes passes max_length - safe_buffer",
    "update_generation_args": "temperature / topp / topk / repetition_penalty overrides of one request",
    # feeding the prompt: False = would exceed max_length - 2 * safe_buffer (a return value, not an exception)
    "prefill": "text -> tokens -> _prefill",
    "append": "text -> tokens (BOS dropped) -> _append",
    "_prefill": "token ids [1, P]: run both models over the prompt, commit the first target token",
    "_append": "token ids [1, A] of a follow-up turn",
    # the loop
    "build_tree": "draft side: grow / fill the token tree behind the committed prefix",
    "verify": "target side: one forward over the tree, accept, compact both caches; False = stop (EOS)",
    # request-level drivers
    "speculative_decoding": "interactive loop -> (decoded tokens, seconds, target steps)",
    "generate": "one request: api_args in, api_args + generated_text / tokens / statistics out",
    "generate_stream": "the same as a generator of (partial_text, perf_log)",
}


class BaseEngine:
    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        missing = [name for name in ENGINE_CONTRACT if not callable(getattr(cls, name, None))]
        if missing:
            raise TypeError(f"{cls.__name__} does not implement the engine contract: missing {missing}")

"""Host helpers of the engines (mirror of the parts of reference
umbrella/speculation/speculation_utils.py the hot path and the front-ends import)."""
from __future__ import annotations

import re
import time

import torch


def make_causal_mask(input_ids_shape, device):
    """speculation_utils.py:4-14: bool [T, T], mask[i, j] = (j <= i)."""
    _, tgt_len = input_ids_shape
    idx = torch.arange(tgt_len, device=device)
    return idx[None, :] < (idx + 1)[:, None]


def find_first_element_position(tensor, elements):
    """speculation_utils.py:316-338."""
    if not isinstance(tensor, torch.Tensor):
        raise ValueError("Input 'tensor' must be a PyTorch Tensor.")
    if not all(isinstance(e, int) for e in elements):
        raise ValueError("All elements in 'elements' must be integers.")
    for idx, value in enumerate(tensor.flatten().tolist()):
        if value in elements:
            return idx
    return -1


def apply_repetition_penalty(input_ids: torch.LongTensor, logits: torch.FloatTensor, penalty: float):
    """speculation_utils.py:340-345."""
    logit = torch.gather(logits, 1, input_ids)
    logit = torch.where(logit < 0, logit * penalty, logit / penalty)
    return logits.scatter(1, input_ids, logit)


def apply_topk(logits: torch.FloatTensor, topk: int):
    """speculation_utils.py:347-352."""
    top_k = min(topk, logits.size(-1))
    remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
    return logits.masked_fill(remove, torch.finfo(logits.dtype).min)


def is_sentence_complete_regex(text):
    """speculation_utils.py:356-358."""
    return bool(re.search(r'[.?!。？！]\s*$', text))


def sample_topk_topp_torch(logits: torch.Tensor, temperature: float, topk: int, topp: float, uniforms: torch.Tensor,
                           static: bool, penalty_ids: torch.Tensor | None = None, penalty_count: int = 0,
                           penalty: float = 1.0) -> torch.Tensor:
    """The sampling of umb_sample_f32 as torch device ops, for top-k values beyond the kernel's shared-memory
    survivor list (topk > 1024; the shipped configs use 32).  Same algorithm: penalty -> top-k (ties at the k-th
    value stay) -> softmax(T) -> nucleus {j: mass(p > p_j) < topp}; ``static``: flashinfer's rejection rounds with
    the fixed uniforms [R, N] (static...:310), else the inverse CDF at uniforms [N] (dynamic...:278-281)."""
    if penalty_ids is not None and penalty > 1.01:
        ids = penalty_ids[:penalty_count].unsqueeze(0).expand(logits.shape[0], -1)
        logits = apply_repetition_penalty(ids, logits, penalty)
    x = logits / temperature if static else logits
    kth = torch.topk(x, min(topk, x.size(-1)))[0][..., -1, None]
    x = x.masked_fill(x < kth, -torch.inf)
    p = torch.softmax(x if static else x / temperature, dim=-1)
    rows = torch.arange(p.shape[0], device=p.device)
    if not static:
        sp = p.sort(dim=-1, descending=True).values
        before = sp.cumsum(-1) - sp                                     # mass strictly ahead in sorted order
        first = torch.ones_like(sp, dtype=torch.bool)
        first[:, 1:] = sp[:, 1:] != sp[:, :-1]
        gt = torch.where(first, before, torch.zeros_like(before)).cummax(-1).values     # ties share their first's mass
        vmin = torch.where(gt < topp, sp, torch.full_like(sp, torch.inf)).min(-1, keepdim=True).values
        kept = torch.where(p >= vmin, p, torch.zeros_like(p))
        cdf = kept.cumsum(-1)
        u = uniforms.reshape(-1, 1) * cdf[:, -1:]
        return (cdf > u).int().argmax(-1)
    q = torch.ones(p.shape[0], 1, device=p.device)
    pivot = torch.zeros_like(q)
    done = torch.zeros(p.shape[0], dtype=torch.bool, device=p.device)
    out = torch.zeros(p.shape[0], dtype=torch.int64, device=p.device)
    for r in range(uniforms.shape[0]):
        live = torch.where(p > pivot, p, torch.zeros_like(p))
        cdf = live.cumsum(-1)
        s = (cdf > uniforms[r].reshape(-1, 1) * q).int().argmax(-1)
        out = torch.where(done, out, s)
        pivot = torch.where(done[:, None], pivot, torch.maximum(pivot, p[rows, s][:, None]))
        q = torch.where(p > pivot, p, torch.zeros_like(p)).sum(-1, keepdim=True)
        done = done | (q[:, 0] < topp)
    return out


class WordStream:
    """Word-wise assembly of streamed text, as the reference's interactive front-ends do it
    (static_speculation_engine.py:487-538, dynamic...:455-507): after every verify the whole answer so far is decoded
    and split on spaces; every word but the last (which the next tokens may still extend) is released, and ``partial``
    is the released text so far.  ``finish`` releases the tail."""

    def __init__(self):
        self.words = [""]
        self.released = 0
        self.partial = ""

    def push(self, decoded_text: str) -> bool:
        """Take the decoded answer so far; True when new complete words were released."""
        self.words = decoded_text.strip().split(" ")
        complete = len(self.words) - 1
        if complete <= self.released:
            return False
        self.partial += " ".join(self.words[self.released:complete]) + " "
        self.released = complete
        return True

    def finish(self) -> str:
        tail = " ".join(self.words[self.released:])
        self.released = len(self.words)
        if tail:
            self.partial += tail
        return self.partial


def stream_generation(engine, max_new_tokens: int, sync=None):
    """The decode loop of ``generate_stream`` for either engine (static...:470-551): yields
    ``(partial_text, perf_log)`` whenever a verify completed new words and once more at the end with the tail.

    Stops like the reference: on the engine's own stop (EOS, ``verify() -> False``), when ``validate_status`` fails,
    at ``max(max_new_tokens, stop_distance)`` tokens, or from ``stop_distance`` tokens before that as soon as the text
    ends a sentence.  ``engine`` is prefilled; it is reset at the end."""
    sync = sync or (lambda: None)
    limit = max(max_new_tokens, engine.stop_distance)
    sync()
    t1 = time.time()
    steps, decode, start = 0, True, engine.num_nodes
    generated_ids: list[int] = []
    stream = WordStream()

    def perf_log() -> str:
        dec_len = engine.num_nodes - start + 1
        return "Output Tokens {} | Avg Accept Tokens {:.2f} | TPOT {:.2f} ms ".format(
            dec_len, dec_len / max(steps, 1), 1000 * (time.time() - t1) / dec_len)

    while decode and engine.validate_status():
        begin = engine.num_nodes
        engine.build_tree()
        decode = engine.verify()
        steps += 1
        generated_ids.extend(engine.tokens[0, begin:engine.num_nodes].tolist())
        if stream.push(engine._decode(generated_ids)):
            yield stream.partial, perf_log()
        done = engine.num_nodes - start
        if done >= limit or (done >= limit - engine.stop_distance and is_sentence_complete_regex(stream.words[-1])):
            decode = False
    stream.finish()
    sync()
    yield stream.partial, perf_log()
    engine.reset()


_TARGET_ONLY_KEYS = ("kv_host_layers", "kv_staging_layers")


def split_model_config(config: dict, target_offload: bool) -> tuple[dict, dict, bool]:
    """The engines hand their whole leftover config to both models' ``alloc`` (static...:92-99).  Keys that ask for
    host-resident KV pages (north_star: 32k contexts over an offloaded KV cache; models/kv_paging.py) concern the
    target only, and need its paged class: returns (draft alloc kwargs, target alloc kwargs, build the target with
    ``offload=True``)."""
    kv_host = int(config.get("kv_host_layers", 0) or 0)
    draft_cfg = {k: v for k, v in config.items() if k not in _TARGET_ONLY_KEYS}
    return draft_cfg, dict(config), bool(target_offload or kv_host > 0)

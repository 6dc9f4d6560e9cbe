"""DynamicSpeculationEngine: SpecExec-style dynamic-tree speculative decoding on the sm_100a verify path.

Host-side mirror of reference umbrella/speculation/dynamic_speculation_engine.py (same constructor
kwargs ``width / num_beams / depth``, same BaseEngine surface, same bookkeeping), SURVEY.md 8a row a2:

  * build_tree() (dynamic...:215-248): ``depth + 1`` draft forwards (1 row, then ``width`` rows per
    level); after each one the draft's top-``num_beams`` logits go through ONE kernel
    (umb_dyn_tree_expand) that scores the beams, keeps the ``width`` best in torch.topk order and
    writes the next level's tokens, parents, scores and ancestor-bit rows.  The reference's
    ``bool[S, S]`` master mask and its per-level row copies do not exist: a node's mask row is
    ``ceil(N/32)`` words (child = parent row | own bit).
  * verify() (dynamic...:250-327): one native target forward over the N = width*depth+1 tree rows
    (chunks of <= 128 rows inside the C call), arg-max fused into the lm_head epilogue for greedy
    decoding, then the same accept / gather kernels as the static engine.  The path test of the
    reference (``rowsum(accept * tree_mask) == depth + 1``) is the parent-chain walk of
    umb_tree_accept.
"""
from __future__ import annotations

import time

import torch

from .. import ops
from ..models.auto_model import AutoModelLM
from .base import BaseEngine
from .speculation_utils import (is_sentence_complete_regex, sample_topk_topp_torch, split_model_config,
                                stream_generation)


class DynamicSpeculationEngine(BaseEngine):

    def __init__(self, draft_model_name: str, target_model_name: str, dtype=torch.float16, device: str = "cuda:0",
                 **kwargs) -> None:
        self.draft_model_name = draft_model_name
        self.target_model_name = target_model_name
        self.dtype = dtype
        self.device = device
        self.max_length = kwargs.pop("max_length", 8192)
        self.stop_distance = kwargs.pop("stop_distance", 32)
        self.safe_buffer = kwargs.pop("safe_buffer", 64)
        self.temperature = kwargs.pop("temperature", 0.0)
        self.topp = kwargs.pop("topp", 0.9)
        self.repetition_penalty = kwargs.pop("repetition_penalty", 1.0)
        self.topk = kwargs.pop("topk", 32)
        self.num_beams = kwargs.pop("num_beams", 24)
        self.tree_width = kwargs.pop("width", 16)
        self.tree_depth = kwargs.pop("depth", 24)
        self.config = kwargs
        self.acceptance_script = None      # benchmark hook: callable(engine, sampled) -> sampled
        self.draft_model = None
        self.target_model = None
        self.stats = {"verify_calls": 0}

    # ------------------------------------------------------------------ set-up (dynamic...:47-88)
    def initialize(self, draft_model=None, target_model=None):
        dev = self.device
        self.storage_ids = torch.arange(self.max_length, device=dev)
        self.tokens = torch.zeros(1, self.max_length, dtype=torch.int64, device=dev)
        self.position_ids = torch.zeros(1, self.max_length, dtype=torch.int64, device=dev)
        self.tree_size = self.tree_width * self.tree_depth + 1
        N = self.tree_size
        assert N < 1024, "tree of more than 1023 nodes"
        self.tree_score = torch.zeros(N, dtype=torch.float32, device=dev)
        self.parents = torch.zeros(N, dtype=torch.int32, device=dev)
        depth = [0]
        for i in range(self.tree_depth):
            depth.extend([i + 1] * self.tree_width)
        self.depth = torch.tensor(depth, dtype=torch.int64, device=dev)         # reference dtype (dynamic...:64)
        self.depth32 = self.depth.to(torch.int32)
        # ancestor-or-self bit rows [N, ceil(N/32)]; row 0 = {root}; rows of level d are rebuilt by build_tree
        self.tree_words = (N + 31) // 32
        self.tree_bits = torch.zeros((N, self.tree_words), dtype=torch.int32, device=dev)
        self.tree_bits[0, 0] = 1

        kw = dict(batch_size=1, max_length=self.max_length, device=dev, dtype=self.dtype)
        draft_cfg, target_cfg, target_offload = split_model_config(self.config, target_offload=True)   # dynamic...:70-80
        if draft_model is None:
            draft_model = AutoModelLM.from_pretrained(model_name=self.draft_model_name, offload=False, **kw)
            draft_model.alloc(**draft_cfg)
        if target_model is None:
            target_model = AutoModelLM.from_pretrained(model_name=self.target_model_name, offload=target_offload, **kw)
            target_model.alloc(**target_cfg)
        self.draft_model, self.target_model = draft_model, target_model
        try:   # tokenizer / generation config need local HF files; token-id entry points work without
            from transformers import AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(self.target_model_name, local_files_only=True)
        except Exception:
            self.tokenizer = None
        self.vocab_size = self.target_model.config.vocab_size
        self.num_nodes = 0
        self._argmax_keys = torch.zeros(N, dtype=torch.int64, device=dev)
        self._result = torch.zeros(4 + N, dtype=torch.int32, device=dev)
        self._result_host = torch.zeros(4 + N, dtype=torch.int32).pin_memory()
        self._set_eos(list(self.target_model.eos_tokens))

    def _set_eos(self, eos_tokens):
        self.eos_tokens = list(eos_tokens)
        self._eos_dev = torch.tensor(self.eos_tokens, dtype=torch.int64, device=self.device) if self.eos_tokens else None

    def _has_room(self, num_nodes: int) -> bool:
        """width * depth + 1 tree rows (385 / 505 in the shipped configs) are written past num_nodes into tokens,
        position_ids and both KV caches -- far more than safe_buffer (64), which is all the reference guards with
        (dynamic...:103, 329); a state without room for the tree is refused instead of reaching the kernels."""
        return num_nodes + self.tree_size + self.tree_depth + 1 <= self.max_length

    # ------------------------------------------------------------------ prefill / append (dynamic...:90-166)
    def _encode(self, text):
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer available offline: use _prefill/_append/generate(input_ids=...)")
        return self.tokenizer.encode(text=text, return_tensors="pt").to(device=self.device)

    def prefill(self, text: str):
        return self._prefill(input_ids=self._encode(text))

    def append(self, text: str):
        return self._append(self._encode(text)[:, 1:])

    def _first_token(self, logits: torch.Tensor):
        """dynamic...:130-133 / :163-165: the first generated token is never an EOS token."""
        if self.eos_tokens:
            logits[-1:, self.eos_tokens] = -torch.inf
        self.tokens[0, self.num_nodes] = ops.argmax(logits[-1:].contiguous())[0]

    @torch.inference_mode()
    def _prefill(self, input_ids: torch.LongTensor):
        prefix_len = input_ids.shape[1]
        if prefix_len >= self.max_length - 2 * self.safe_buffer or not self._has_room(self.num_nodes + prefix_len):
            return False
        self.num_nodes += prefix_len
        self.num_nodes_this_iter = self.num_nodes + self.tree_size
        self.num_draft_model_tokens_this_iter = self.num_nodes
        self.tokens[:, :prefix_len].copy_(input_ids)
        self.position_ids[:, :prefix_len] = torch.arange(prefix_len, device=self.device).unsqueeze(0)
        self.position_ids[:, prefix_len:prefix_len + self.tree_size] = prefix_len + self.depth
        ids, pos = self.tokens[0, :prefix_len], self.position_ids[0, :prefix_len]
        self.draft_model.prefill_forward(ids, pos, want_logits=False)
        self._first_token(self.target_model.prefill_forward(ids, pos, last_only=True))
        return True

    @torch.inference_mode()
    def _append(self, input_ids: torch.LongTensor):
        append_len = input_ids.shape[1]
        if append_len + self.num_nodes >= self.max_length - 2 * self.safe_buffer \
                or not self._has_room(self.num_nodes + append_len + 1):
            return False
        self.tokens[:, self.num_nodes + 1:self.num_nodes + 1 + append_len].copy_(input_ids)
        last = self.num_nodes
        self.num_nodes += append_len + 1
        self.num_nodes_this_iter = self.num_nodes + self.tree_size
        self.num_draft_model_tokens_this_iter = self.num_nodes
        self.position_ids[:, :self.num_nodes] = torch.arange(self.num_nodes, device=self.device).unsqueeze(0)
        self.position_ids[:, self.num_nodes:self.num_nodes + self.tree_size] = self.num_nodes + self.depth
        ids, pos = self.tokens[0, last:self.num_nodes], self.position_ids[0, last:self.num_nodes]
        self.draft_model.prefill_forward(ids, pos, want_logits=False)
        self._first_token(self.target_model.prefill_forward(ids, pos, last_only=True))
        return True

    # ------------------------------------------------------------------ the loop
    @torch.inference_mode()
    def build_tree(self):
        """dynamic...:215-248.  Level 0 is the root alone, level d >= 1 holds ``width`` nodes at tree
        ids [1 + (d-1) width, 1 + d width)."""
        n = self.num_nodes
        for step in range(self.tree_depth + 1):
            dec_len = self.tree_width if step > 0 else 1
            first = self.num_draft_model_tokens_this_iter - n          # tree id of the level's first node
            a, b = n + first, n + first + dec_len
            last = step == self.tree_depth
            logits = self.draft_model.verify_forward(self.tokens[0, a:b], self.position_ids[0, a:b], n,
                                                     self.tree_bits[first:first + dec_len], want_logits=not last)
            self.num_draft_model_tokens_this_iter += dec_len
            if not last:
                if self.num_beams <= ops.TOPK_MAX:
                    top_logits, token_ids = ops.topk(logits, self.num_beams)       # one launch instead of torch's five
                else:
                    top_logits, token_ids = logits.topk(dim=-1, k=self.num_beams)
                ops.dyn_tree_expand(top_logits, token_ids, self.tree_width, first, n, self.tree_score, self.tokens[0],
                                    self.parents, self.tree_bits)

    def _sample_target(self, logits: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
        """dynamic...:265-281.  Greedy: decode the arg-max keys produced by the lm_head epilogue.  Otherwise ONE fused
        kernel (umb_sample_f32): repetition penalty, apply_topk, softmax(T), top_p_renorm_prob's nucleus and the draw.
        The reference draws with torch.multinomial; here one uniform per node from torch's generator (so
        torch.manual_seed reproduces a run) goes through the inverse CDF -- the same distribution."""
        n, N = self.num_nodes, self.tree_size
        penalised = self.repetition_penalty > 1.01
        if self.temperature < 0.05 and not penalised:
            return ops.argmax_keys_to_ids(keys)       # also resets the keys for the next forward
        ops.argmax_keys_to_ids(keys)
        pen = dict(penalty_ids=self.tokens[0], penalty_count=n + 1, penalty=self.repetition_penalty) if penalised else {}
        if self.temperature < 0.05:
            return ops.sample(logits, ops.UMB_SAMPLE_ARGMAX, **pen)
        self.last_uniforms = torch.rand(N, device=logits.device)
        if self.topk > ops.SAMPLE_MAX_TOPK:
            return sample_topk_topp_torch(logits, self.temperature, self.topk, self.topp, self.last_uniforms,
                                          static=False, **pen)
        return ops.sample(logits, ops.UMB_SAMPLE_DYNAMIC, temperature=self.temperature, top_k=self.topk, top_p=self.topp,
                          uniforms=self.last_uniforms, **pen)

    @torch.inference_mode()
    def verify(self):
        n, N = self.num_nodes, self.tree_size
        ids = self.tokens[0, n:n + N]
        pos = self.position_ids[0, n:n + N]
        logits = self.target_model.verify_forward(ids, pos, n, self.tree_bits, argmax_keys=self._argmax_keys)
        sampled = self._sample_target(logits, self._argmax_keys[:N])
        if self.acceptance_script is not None:
            sampled = self.acceptance_script(self, sampled)
        self.last_sampled = sampled
        ops.tree_accept(sampled, self.tokens[0], self.position_ids[0], self.parents, self.depth32, n, self._eos_dev,
                        self._result)
        self.draft_model.kv_cache.gather_from_result(self._result, n, 0)
        self.target_model.kv_cache.gather_from_result(self._result, n, 0)
        self._result_host.copy_(self._result, non_blocking=True)
        self.parents.zero_()                     # dynamic...:325-326
        self.tree_score.zero_()
        torch.cuda.current_stream().synchronize()
        accept_length, cont, new_n = (int(x) for x in self._result_host[:3])
        self.draft_model.kv_cache.kv_offset = n + accept_length
        self.target_model.kv_cache.kv_offset = n + accept_length
        self.num_nodes = new_n
        self.num_draft_model_tokens_this_iter = new_n
        self.num_nodes_this_iter = new_n + N
        self.stats["verify_calls"] += 1
        self.last_accept_length = accept_length
        return bool(cont)

    @torch.inference_mode()
    def reset(self):
        self.num_nodes = 0
        self.parents.zero_()
        self.tree_score.zero_()
        self.tokens.zero_()
        self.position_ids.zero_()
        self.draft_model.clear()
        self.target_model.clear()

    def validate_status(self):
        return self.num_nodes <= (self.max_length - self.safe_buffer) and self._has_room(self.num_nodes)

    def update_generation_args(self, **generation_args):
        self.temperature = generation_args.pop("temperature", self.temperature)
        self.topp = generation_args.pop("topp", self.topp)
        self.repetition_penalty = generation_args.pop("repetition_penalty", self.repetition_penalty)
        self.topk = generation_args.pop("topk", self.topk)

    def _decode(self, ids):
        if self.tokenizer is None:
            return ""
        return self.tokenizer.decode(ids, skip_special_tokens=True, clean_up_tokenization_spaces=False)

    @torch.inference_mode()
    def speculative_decoding(self, max_new_tokens=128):
        """dynamic...:167-213 -> (dec_len, seconds, large_model_steps)."""
        max_new_tokens = max(max_new_tokens, self.stop_distance)
        torch.cuda.synchronize()
        t1 = time.time()
        steps, decode, start = 0, True, self.num_nodes
        generated_ids, pos = [], 0
        text = [""]
        while decode and self.validate_status():
            begin = self.num_nodes
            self.build_tree()
            decode = self.verify()
            steps += 1
            generated_ids.extend(self.tokens[0, begin:self.num_nodes].tolist())
            text = self._decode(generated_ids).strip().split(" ")
            now = len(text) - 1
            if now > pos:
                print(" ".join(text[pos:now]), end=" ", flush=True)
                pos = now
            done_len = self.num_nodes - start
            if (is_sentence_complete_regex(text[-1]) and done_len >= max_new_tokens - self.stop_distance) \
                    or done_len >= max_new_tokens:
                decode = False
        print(" ".join(text[pos:]), flush=True)
        torch.cuda.synchronize()
        t2 = time.time()
        return self.num_nodes - start + 1, t2 - t1, steps

    def _empty(self, api_args):
        api_args.update(generated_text="", generated_tokens=[], avg_accept_tokens=0, time_per_output_token=0)
        return api_args

    def _start(self, api_args):
        """Common head of generate / generate_stream (dynamic...:346-381): returns None when there is
        nothing to generate, else whether the prefill fitted."""
        self.update_generation_args(**api_args)
        input_ids = api_args.get("input_ids", None)
        max_new_tokens = api_args.get("max_new_tokens", 128)
        if input_ids is None:
            context = api_args.get("context", None)
            if context is None or len(context) == 0 or max_new_tokens == 0:
                return None
            return self.prefill(context)
        if len(input_ids) == 0 or max_new_tokens == 0:
            return None
        ids = torch.as_tensor(input_ids, dtype=torch.int64)
        if not ids.is_cuda:
            ids = ids.pin_memory().to(self.device, non_blocking=True) if torch.cuda.is_available() else ids
        return self._prefill(input_ids=ids.reshape(1, -1))

    @torch.inference_mode()
    def generate(self, **api_args):
        """dynamic...:343-409."""
        success = self._start(api_args)
        if not success:
            if success is not None:
                self.reset()
            return self._empty(api_args)
        max_new_tokens = api_args.get("max_new_tokens", 128)
        torch.cuda.synchronize()
        t1 = time.time()
        steps, decode, start = 0, True, self.num_nodes
        while decode and (self.num_nodes - start) < max_new_tokens and self.validate_status():
            self.build_tree()
            decode = self.verify()
            steps += 1
        torch.cuda.synchronize()
        t2 = time.time()
        dec_len = self.num_nodes - start + 1
        out_ids = self.tokens[0, start:self.num_nodes + 1].tolist()
        api_args["generated_text"] = self._decode(out_ids)
        api_args["generated_tokens"] = out_ids
        api_args["avg_accept_tokens"] = dec_len / max(steps, 1)
        api_args["time_per_output_token"] = 1000 * (t2 - t1) / dec_len
        api_args["large_model_steps"] = steps
        api_args["decode_seconds"] = t2 - t1
        self.reset()
        return api_args

    @torch.inference_mode()
    def generate_stream(self, **api_args):
        """dynamic...:412-: yields (text released so far, perf_log) whenever a verify completed new words, and the whole
        text at the end (``speculation_utils.stream_generation``)."""
        success = self._start(api_args)
        if success is None:                                          # nothing to generate
            yield "", ""
            return
        if not success:
            self.reset()
            yield "Exceeding reserved allowed context length", "Exceeding reserved allowed context length"   # dynamic...:444
            return
        yield from stream_generation(self, api_args.get("max_new_tokens", 128), sync=torch.cuda.synchronize)

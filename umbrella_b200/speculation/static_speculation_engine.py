"""StaticSpeculationEngine: Sequoia static-tree speculative decoding on the sm_100a verify path.

Host-side mirror of reference umbrella/speculation/static_speculation_engine.py (same constructor
kwargs, same 13-method BaseEngine surface, same return values and bookkeeping), with the inner loop
replaced:

  * verify(): one native target forward over the token tree (umb_llama_forward: AWQ tcgen05 GEMMs,
    tree attention with (prefix_len, ancestor bitmask) instead of the O(S^2) master mask, arg-max
    fused into the lm_head epilogue), then ONE accept kernel (umb_tree_accept: accept flags, path
    compaction, token/position writes, EOS truncation, state advance) and ONE gather kernel per
    model (umb_kv_gather); the host reads back 4 ints per iteration from pinned memory.
  * no ``attn_mask`` [S, 2S], no per-step mask ``.contiguous()`` copy (static...:55-57,345), no
    ``nonzero()`` / ``.tolist()`` syncs (:320,:330).
"""
from __future__ import annotations

import time

import torch

from .. import _lib, ops
from ..models.auto_model import AutoModelLM
from ..sequoia_utils import load_growmap, tree_bits_from_mask
from .base import BaseEngine
from .speculation_utils import (is_sentence_complete_regex, sample_topk_topp_torch, split_model_config,
                                stream_generation)


class StaticSpeculationEngine(BaseEngine):

    def __init__(self, draft_model_name: str, target_model_name: str, dtype=torch.float16, device: str = "cuda:0",
                 **kwargs) -> None:
        self.draft_model_name = draft_model_name
        self.target_model_name = target_model_name
        self.dtype = dtype
        self.device = device
        self.growmap_path = kwargs.pop("growmap_path", None)
        assert self.growmap_path is not None, "Please specify growmap path for static trees"
        self.max_length = kwargs.pop("max_length", 8192)
        self.stop_distance = kwargs.pop("stop_distance", 32)
        self.safe_buffer = kwargs.pop("safe_buffer", 64)
        self.temperature = kwargs.pop("temperature", 0.0)
        self.topp = kwargs.pop("topp", 0.9)
        self.repetition_penalty = kwargs.pop("repetition_penalty", 1.0)
        self.topk = kwargs.pop("topk", 32)
        self.config = kwargs
        self.acceptance_script = None      # benchmark hook: callable(engine, sampled) -> sampled
        # One CUDA graph per speculative-decoding iteration (all draft levels + verify + accept + gathers):
        # the reference graphs only the draft forward (llama.py:505-533); here the whole iteration is one
        # launch and the host reads back 3 ints.  Greedy decoding only; other settings run eagerly.
        self.graph_iteration = kwargs.pop("graph_iteration", True)
        self.graph_bucket = kwargs.pop("graph_bucket", 512)
        # Draft-side schedule.  The reference runs one draft forward per tree level, the last one only to put
        # the leaves' K/V into the draft cache (static...:257-281).  With lazy_leaf_kv the leaf level is not run:
        # a leaf's K/V is needed only if that leaf is accepted, and then it is the LAST committed token -- the
        # next iteration's first draft step recomputes it, feeding [last committed token, root] as two causal
        # rows.  One draft forward fewer per iteration; the accepted tokens are unchanged (they are the target's).
        self.lazy_leaf_kv = kwargs.pop("lazy_leaf_kv", True)
        self.draft_model = None
        self.target_model = None
        self.stats = {"verify_calls": 0}

    # ------------------------------------------------------------------ set-up (static...:48-131)
    def initialize(self, draft_model=None, target_model=None):
        dev = self.device
        self.growmap = load_growmap(self.growmap_path)
        self.storage_ids = torch.arange(self.max_length, device=dev)
        self.tokens = torch.zeros(1, self.max_length, dtype=torch.int64, device=dev)
        self.position_ids = torch.zeros(1, self.max_length, dtype=torch.int64, device=dev)

        idx_lists = self.growmap["roots"]
        self.tree_depth = len(idx_lists)
        self.level_ranges = [(lvl[0], lvl[-1] + 1) for lvl in idx_lists]      # levels are contiguous id ranges
        for (a, b), lvl in zip(self.level_ranges, idx_lists):
            assert list(range(a, b)) == list(lvl), "growmap levels must be contiguous node ranges"
        self.Successors = self.growmap["Successors"]
        self.tree_size = self.growmap["size"]
        tree_mask = torch.tensor(self.growmap["mask"]) == 1
        self.tree_mask = tree_mask.to(dev)
        self.node_in_path = self.tree_mask.int().sum(dim=-1)
        self.tree_bits = torch.from_numpy(tree_bits_from_mask(tree_mask.numpy())).to(dev)
        parents = torch.zeros(self.tree_size, dtype=torch.int32)
        for v, succ in enumerate(self.Successors):
            for c in succ:
                parents[c] = v
        self.parents = parents.to(dev)
        self.depth = torch.tensor(self.growmap["depth"], dtype=torch.int32, device=dev)
        self.depth64 = self.depth.to(torch.int64)
        self.branch_lists = self.growmap["branches"]

        kw = dict(batch_size=1, max_length=self.max_length, device=dev, dtype=self.dtype)
        # "kv_host_layers" in the config: the target keeps that many layers' K/V pages in pinned host memory (paged class)
        draft_cfg, target_cfg, target_offload = split_model_config(self.config, target_offload=False)
        if draft_model is None:
            draft_model = AutoModelLM.from_pretrained(model_name=self.draft_model_name, offload=False, cuda_graph=True, **kw)
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
        self.eos_tokens = list(self.target_model.eos_tokens)
        self.num_nodes = 0

        # per-level gather of the draft's top-k samples (static...:115-123)
        self.sample_gather_indices = {}
        for i in range(self.tree_depth - 1):
            width = max(self.branch_lists[i])
            idx = [torch.arange(b, dtype=torch.long) + j * width for j, b in enumerate(self.branch_lists[i])]
            self.sample_gather_indices[i] = torch.cat(idx).to(dev)

        N = self.tree_size
        self._argmax_keys = torch.zeros(max(N, 1), dtype=torch.int64, device=dev)
        self._result = torch.zeros(4 + N, dtype=torch.int32, device=dev)
        self._result_host = torch.zeros(4 + N, dtype=torch.int32).pin_memory()
        self._set_eos(self.eos_tokens)
        self.uniform_samples = torch.rand(3, N, device=dev)
        # device-resident iteration state for graph replay: [n, n + N, n before the last accept, spare]
        self._state = torch.zeros(4, dtype=torch.int32, device=dev)
        self._graphs = {}
        self._tree_pending = False
        V = self.vocab_size
        self._draft_logits = [torch.empty((b - a + (1 if a == 0 else 0), self.draft_model.config.vocab_size),
                                          dtype=torch.float32, device=dev) for (a, b) in self.level_ranges[:-1]]
        self._root_bits = torch.tensor([[1], [3]], dtype=torch.int32, device=dev)      # [prev, root]: causal pair
        self._target_logits = torch.empty((N, V), dtype=torch.float32, device=dev)

    def _set_eos(self, eos_tokens):
        self.eos_tokens = list(eos_tokens)
        self._eos_dev = torch.tensor(self.eos_tokens, dtype=torch.int64, device=self.device) if self.eos_tokens else None
        self._graphs = {}                  # captured graphs hold the old pointer

    def _sync_state(self):
        """Push the host's num_nodes to the device state word (after prefill / append / reset)."""
        self._state[:3] = torch.tensor([self.num_nodes, self.num_nodes + self.tree_size, self.num_nodes],
                                       dtype=torch.int32)

    def _has_room(self, num_nodes: int) -> bool:
        """The kernels index tokens / position_ids / both KV caches at rows up to num_nodes + tree_size (tree rows,
        the next tree's positions, the bonus token).  The reference only guards with safe_buffer and fails with a
        shape error beyond it (static...:146-147, 351); here a state without room is refused (``False``), so a
        growmap larger than safe_buffer or a long multi-turn prefix can never reach a device-side write."""
        return num_nodes + self.tree_size + self.tree_depth + 1 <= self.max_length

    # ------------------------------------------------------------------ prefill / append (static...:133-209)
    def _encode(self, text):
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer available offline: use _prefill/_append/generate(input_ids=...)")
        return self.tokenizer.encode(text=text, return_tensors="pt").to(device=self.device)

    def prefill(self, text: str):
        return self._prefill(input_ids=self._encode(text))

    def append(self, text: str):
        return self._append(self._encode(text)[:, 1:])

    @torch.inference_mode()
    def _prefill(self, input_ids: torch.LongTensor):
        prefix_len = input_ids.shape[1]
        if prefix_len >= self.max_length - 2 * self.safe_buffer or not self._has_room(self.num_nodes + prefix_len):
            return False
        self.num_nodes += prefix_len
        self.num_nodes_this_iter = self.num_nodes + self.tree_size
        self.num_draft_tokens_this_iter = self.num_nodes
        self.tokens[:, :prefix_len].copy_(input_ids)
        self.position_ids[:, :prefix_len] = torch.arange(prefix_len, device=self.device).unsqueeze(0)
        self.position_ids[:, prefix_len:prefix_len + self.tree_size] = prefix_len + self.depth64
        ids, pos = self.tokens[0, :prefix_len], self.position_ids[0, :prefix_len]
        self.draft_model.prefill_forward(ids, pos, want_logits=False)
        logits = self.target_model.prefill_forward(ids, pos, last_only=True)
        self.tokens[0, self.num_nodes] = ops.argmax(logits)[0]
        self._sync_state()
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
        self.num_draft_tokens_this_iter = self.num_nodes
        self.position_ids[:, :self.num_nodes] = torch.arange(self.num_nodes, device=self.device).unsqueeze(0)
        self.position_ids[:, self.num_nodes:self.num_nodes + self.tree_size] = self.num_nodes + self.depth64
        ids, pos = self.tokens[0, last:self.num_nodes], self.position_ids[0, last:self.num_nodes]
        if bool(self.lazy_leaf_kv) and self.tree_depth >= 2 and last >= 1:
            # lazy_leaf_kv leaves the draft K/V row of an accepted last-level node (row last - 1) unwritten and
            # relies on the NEXT build_tree recomputing it; an append comes in between, so the draft prefill starts
            # one row earlier and re-derives that row together with the appended ones
            self.draft_model.kv_cache.kv_offset = last - 1
            self.draft_model.prefill_forward(self.tokens[0, last - 1:self.num_nodes],
                                             self.position_ids[0, last - 1:self.num_nodes], want_logits=False)
        else:
            self.draft_model.prefill_forward(ids, pos, want_logits=False)
        logits = self.target_model.prefill_forward(ids, pos, last_only=True)
        self.tokens[0, self.num_nodes] = ops.argmax(logits)[0]
        self._sync_state()
        return True

    # ------------------------------------------------------------------ the loop
    def _lazy(self) -> bool:
        return bool(self.lazy_leaf_kv) and self.tree_depth >= 2 and self.num_nodes >= 1

    def _graph_ok(self) -> bool:
        return bool(self.graph_iteration) and self.temperature < 0.05 and not (self.repetition_penalty > 1.01) \
            and getattr(self.target_model, "supports_graph", True) and getattr(self.draft_model, "supports_graph", True)

    def _iteration_body(self, bound: int):
        """Enqueue one whole iteration with every position read from the device state word."""
        N = self.tree_size
        tok, pos, st = self.tokens[0], self.position_ids[0], self._state
        lazy = self._lazy()
        for step in range(self.tree_depth - 1 if lazy else self.tree_depth):
            a, b = self.level_ranges[step]
            last = step == self.tree_depth - 1
            lg = None if last else self._draft_logits[step][:b - a]
            if lazy and step == 0:
                self.draft_model.forward_dyn(tok, pos, 2, st, -1, bound, self._root_bits, logits=self._draft_logits[0])
                lg = self._draft_logits[0][1:2]
            else:
                self.draft_model.forward_dyn(tok, pos, b - a, st, a, bound, self.tree_bits[a:b], logits=lg)
            if not last:
                self._next_level_tokens(step, lg, b, n_dev=st)
        self.target_model.forward_dyn(tok, pos, N, st, 0, bound, self.tree_bits, logits=self._target_logits,
                                      argmax_keys=self._argmax_keys)
        sampled = ops.argmax_keys_to_ids(self._argmax_keys[:N])
        if self.acceptance_script is not None:
            sampled = self.acceptance_script.graph_call(self, sampled, st)
        ops.tree_accept(sampled, tok, pos, self.parents, self.depth, 0, self._eos_dev, self._result, state=st)
        self.draft_model.kv_cache.gather_from_result(self._result, 0, 0, offset_dev=st[2:3])
        self.target_model.kv_cache.gather_from_result(self._result, 0, 0, offset_dev=st[2:3])
        self._result_host.copy_(self._result, non_blocking=True)

    def _run_iteration_graph(self):
        n, N = self.num_nodes, self.tree_size
        bucket = self.graph_bucket
        bound = min(self.max_length, ((n + N) // bucket + 1) * bucket)
        key = (bound, id(self.acceptance_script), bool(self._lazy()))
        g = self._graphs.get(key)
        if g is None:
            # first iteration of this bucket runs eagerly (it is also the warm-up), then the graph is captured
            self._iteration_body(bound)
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            script_it = None if self.acceptance_script is None else self.acceptance_script.snapshot()
            before = _lib.launch_count()
            with torch.cuda.graph(graph):
                self._iteration_body(bound)
            graph_launches = _lib.launch_count() - before          # kernels one replay re-issues
            _lib.load().umb_launch_count_add(-graph_launches & 0xFFFFFFFFFFFFFFFF)   # capture itself ran nothing
            if script_it is not None:
                self.acceptance_script.restore(script_it)
            self._graphs[key] = (graph, graph_launches)
        else:
            g[0].replay()
            _lib.load().umb_launch_count_add(g[1])
            torch.cuda.current_stream().synchronize()
        accept_length, cont, new_n = (int(x) for x in self._result_host[:3])
        self.draft_model.kv_cache.kv_offset = n + accept_length
        self.target_model.kv_cache.kv_offset = n + accept_length
        self.num_nodes = new_n
        self.num_draft_tokens_this_iter = new_n
        self.num_nodes_this_iter = new_n + N
        self.stats["verify_calls"] += 1
        self.last_accept_length = accept_length
        return bool(cont)

    def _next_level_tokens(self, step: int, logits: torch.Tensor, b: int, num_nodes: int = 0, n_dev=None):
        """sampling_argmax_gather (speculation_utils.py:52-61; static...:274-281): the per-node top-``max(branches)``
        draft tokens, picked in growmap order, become tree nodes [b, b + total).  Two launches: umb_topk_f32 and the
        gathering token store (torch: five mbtopk kernels + flatten + index_select + copy)."""
        width = max(self.branch_lists[step])
        gather = self.sample_gather_indices[step]
        if width <= ops.TOPK_MAX:
            _, ids = ops.topk(logits, width)
            ops.store_tokens_gather(self.tokens[0], ids, gather, b, num_nodes=num_nodes, n_dev=n_dev)
            return
        new_tokens = logits.topk(k=width).indices.flatten().index_select(0, gather)
        if n_dev is not None:
            ops.store_tokens(self.tokens[0], n_dev, b, new_tokens)
        else:
            self.tokens[0, num_nodes + b:num_nodes + b + new_tokens.numel()] = new_tokens

    @torch.inference_mode()
    def build_tree(self):
        """Draft side (static...:257-281): one draft forward per tree level; the next level's tokens
        are the per-node top-k draft tokens gathered in growmap order."""
        if self._graph_ok():
            self._tree_pending = True          # the draft levels run inside verify()'s iteration graph
            return
        n = self.num_nodes
        lazy = self._lazy()
        for step in range(self.tree_depth - 1 if lazy else self.tree_depth):
            a, b = self.level_ranges[step]
            last = step == self.tree_depth - 1
            if lazy and step == 0:
                self.draft_model.kv_cache.kv_offset = n - 1
                logits = self.draft_model.verify_forward(self.tokens[0, n - 1:n + 1], self.position_ids[0, n - 1:n + 1],
                                                         n - 1, self._root_bits)[1:]
            else:
                ids = self.tokens[0, n + a:n + b]
                pos = self.position_ids[0, n + a:n + b]
                logits = self.draft_model.verify_forward(ids, pos, n, self.tree_bits[a:b], want_logits=not last)
            self.num_draft_tokens_this_iter += b - a
            if not last:
                self._next_level_tokens(step, logits, b, num_nodes=n)
        if lazy:       # leaf rows stay unwritten; the gather treats the cache as holding all N tree rows
            self.draft_model.kv_cache.kv_offset = n + self.tree_size
            self.num_draft_tokens_this_iter = n + self.tree_size

    def _sample_target(self, logits: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
        """static...:298-310.  Greedy: decode the arg-max keys produced by the lm_head epilogue.  Otherwise ONE fused
        kernel (umb_sample_f32): repetition penalty over tokens[:n + 1], top-k, softmax(T), and the top-p rejection
        rounds of flashinfer's top_k_top_p_sampling_from_logits driven by the fixed ``uniform_samples`` [3, N] drawn
        at initialize() (static...:131) -- like the reference, a deterministic function of the logits."""
        n = self.num_nodes
        penalised = self.repetition_penalty > 1.01
        if self.temperature < 0.05 and not penalised:
            return ops.argmax_keys_to_ids(keys)
        ops.argmax_keys_to_ids(keys)                  # drop (and reset) the un-penalised keys
        pen = dict(penalty_ids=self.tokens[0], penalty_count=n + 1, penalty=self.repetition_penalty) if penalised else {}
        if self.temperature < 0.05:
            return ops.sample(logits, ops.UMB_SAMPLE_ARGMAX, **pen)
        if self.topk > ops.SAMPLE_MAX_TOPK:
            return sample_topk_topp_torch(logits, self.temperature, self.topk, self.topp, self.uniform_samples,
                                          static=True, **pen)
        return ops.sample(logits, ops.UMB_SAMPLE_STATIC, temperature=self.temperature, top_k=self.topk, top_p=self.topp,
                          uniforms=self.uniform_samples, **pen)

    @torch.inference_mode()
    def verify(self):
        if self._tree_pending:
            self._tree_pending = False
            return self._run_iteration_graph()
        n, N = self.num_nodes, self.tree_size
        ids = self.tokens[0, n:n + N]
        pos = self.position_ids[0, n:n + N]
        logits = self.target_model.verify_forward(ids, pos, n, self.tree_bits, argmax_keys=self._argmax_keys)
        sampled = self._sample_target(logits, self._argmax_keys[:N])
        if self.acceptance_script is not None:
            sampled = self.acceptance_script(self, sampled)
        ops.tree_accept(sampled, self.tokens[0], self.position_ids[0], self.parents, self.depth, n, self._eos_dev,
                        self._result, state=self._state)
        # compaction reads (len, path) from the device result; the host only needs 4 ints afterwards
        self.draft_model.kv_cache.gather_from_result(self._result, n, 0)
        self.target_model.kv_cache.gather_from_result(self._result, n, 0)
        self._result_host.copy_(self._result, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        accept_length, cont, new_n = (int(x) for x in self._result_host[:3])
        self.draft_model.kv_cache.kv_offset = n + accept_length
        self.target_model.kv_cache.kv_offset = n + accept_length
        self.num_nodes = new_n
        self.num_draft_tokens_this_iter = new_n
        self.num_nodes_this_iter = new_n + N
        self.stats["verify_calls"] += 1
        self.last_accept_length = accept_length
        return bool(cont)

    @torch.inference_mode()
    def reset(self):
        self.num_nodes = 0
        self._tree_pending = False
        self.tokens.zero_()
        self.position_ids.zero_()
        self.draft_model.clear()
        self.target_model.clear()
        self._state.zero_()

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
        """static...:211-255 -> (dec_len, seconds, large_model_steps)."""
        max_new_tokens = max(max_new_tokens, self.stop_distance)
        torch.cuda.synchronize()
        t1 = time.time()
        steps, decode, start = 0, True, self.num_nodes
        generated_ids, pos = [], 0
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
        print(" ".join(text[pos:]) if steps else "", flush=True)   # the last words (static...:252)
        torch.cuda.synchronize()
        t2 = time.time()
        return self.num_nodes - start + 1, t2 - t1, steps

    def _empty(self, api_args):
        api_args.update(generated_text="", generated_tokens=[], avg_accept_tokens=0, time_per_output_token=0)
        return api_args

    @torch.inference_mode()
    def generate(self, **api_args):
        """static...:373-434: prefill (text or token ids), decode loop, stats, reset."""
        self.update_generation_args(**api_args)
        input_ids = api_args.get("input_ids", None)
        max_new_tokens = api_args.get("max_new_tokens", 128)
        if input_ids is None:
            context = api_args.get("context", None)
            if context is None or len(context) == 0 or max_new_tokens == 0:
                return self._empty(api_args)
            success = self.prefill(context)
        else:
            if len(input_ids) == 0 or max_new_tokens == 0:
                return self._empty(api_args)
            ids = torch.as_tensor(input_ids, dtype=torch.int64)
            if not ids.is_cuda:
                ids = ids.pin_memory().to(self.device, non_blocking=True) if torch.cuda.is_available() else ids
            success = self._prefill(input_ids=ids.reshape(1, -1))
        if not success:
            self.reset()
            return self._empty(api_args)
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
        """static...:437-551: yields (text released so far, perf_log) whenever a verify completed new words, and the
        whole text at the end (``speculation_utils.stream_generation``)."""
        self.update_generation_args(**api_args)
        input_ids = api_args.get("input_ids", None)
        max_new_tokens = api_args.get("max_new_tokens", 128)
        if input_ids is None:
            context = api_args.get("context", None)
            if not context or max_new_tokens == 0:
                yield "", ""
                return
            success = self.prefill(context)
        else:
            if len(input_ids) == 0 or max_new_tokens == 0:          # static...:459-464
                yield "", ""
                return
            ids = torch.as_tensor(input_ids, dtype=torch.int64).to(self.device)
            success = self._prefill(input_ids=ids.reshape(1, -1))
        if not success:
            self.reset()
            yield "Exceeding reserved allowed context length", "Exceeding reserved allowed context length"   # static...:469
            return
        yield from stream_generation(self, max_new_tokens, sync=torch.cuda.synchronize)

"""KV cache + tree-masked attention seam (host-side mirror of reference umbrella/attn/cache.py:5-96).

Same public surface as the reference's ``KV_Cache`` -- ``update_kv_cache, compute_attention,
gather_kv_incremental, clear, set_kv_len`` and the ``k_cache / v_cache / kv_offset`` attributes with
the same ``[L, S, Hkv, D]`` layout -- with the three GPU operations routed to the sm_100a kernels:
append (``update_kv_cache``), masked attention (``compute_attention`` -> umb_tree_attn_f16 instead of
flashinfer.single_prefill_with_kv_cache) and accept-path compaction (``gather_kv_incremental`` ->
umb_kv_gather).
"""
from __future__ import annotations

import torch

from .. import ops


class KV_Cache:
    def __init__(self, config, batch_size: int = 1, max_length: int = 256, device="cuda:0", dtype=torch.float16,
                 num_layers: int | None = None, zero_tail: bool = False) -> None:
        assert batch_size == 1 and dtype == torch.float16
        self.config = config
        self.max_length = max_length
        self.device = torch.device(device)
        self.dtype = dtype
        self.head_dim = getattr(config, "head_dim", config.hidden_size // config.num_attention_heads)
        self.num_layers = config.num_hidden_layers if num_layers is None else num_layers
        shape = (self.num_layers, max_length, config.num_key_value_heads, self.head_dim)
        self.k_cache = torch.zeros(shape, device=self.device, dtype=dtype)
        self.v_cache = torch.zeros(shape, device=self.device, dtype=dtype)
        self.kv_offset = 0
        self.num_key_value_heads = config.num_key_value_heads
        self.num_attention_heads = config.num_attention_heads
        self.num_key_value_groups = config.num_attention_heads // config.num_key_value_heads
        # The reference zero-fills [offset + m, S) of every layer after each gather (cache.py:46-47):
        # up to the whole cache per step, and dead work (only [:kv_offset] is ever read).  Off by
        # default; tests switch it on for byte-for-byte cache comparison.
        self.zero_tail = zero_tail
        self._gather_buf = torch.zeros(4 + 1024, dtype=torch.int32, device=self.device)
        self._attn_ws = None

    # cache.py:41-49 -- indices: absolute rows, ascending; rows land at [offset, offset + len)
    def gather_kv_incremental(self, indices: torch.Tensor, offset: int):
        m = int(indices.numel())
        buf = self._gather_buf
        if m > buf.numel() - 4:
            raise ValueError("accept path longer than 1024 rows")
        buf[0] = m
        if m:
            buf[4:4 + m] = indices.to(device=self.device, dtype=torch.int32)
        self.gather_from_result(buf, offset, m)

    def gather_from_result(self, result: torch.Tensor, offset: int, accept_len: int, offset_dev=None):
        """Fast path: ``result`` is the device buffer written by umb_tree_accept (no host round trip
        for the path itself); ``offset_dev`` (device int32) overrides ``offset`` for graph replay."""
        ops.kv_gather(self.k_cache, self.v_cache, result, offset, "LSHD",
                      zero_upto=self.max_length if self.zero_tail else 0, offset_dev=offset_dev)
        self.kv_offset = offset + accept_len

    # cache.py:53-65 -- append at kv_offset; the values of storage_ids are ignored, only their count
    def update_kv_cache(self, new_k_cache: torch.Tensor, new_v_cache: torch.Tensor, layer_idx: int,
                        storage_ids: torch.Tensor):
        n = storage_ids.shape[0]
        if layer_idx == 0:
            self.kv_offset += n
        self.k_cache[layer_idx][self.kv_offset - n:self.kv_offset] = new_k_cache
        self.v_cache[layer_idx][self.kv_offset - n:self.kv_offset] = new_v_cache
        return self.k_cache[layer_idx][:self.kv_offset], self.v_cache[layer_idx][:self.kv_offset]

    # cache.py:67-87
    def compute_attention(self, query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor,
                          layer_idx, storage_ids: torch.Tensor, attention_mask: torch.Tensor, logits_soft_cap=0):
        self.update_kv_cache(key_states[0], value_states[0], layer_idx, storage_ids)
        mask = attention_mask if attention_mask.dtype in (torch.bool, torch.uint8) else attention_mask != 0
        if mask.stride(-1) != 1:
            mask = mask.contiguous()
        q = query_states[0].contiguous()
        if self._attn_ws is None or self._attn_ws_rows < q.shape[0]:
            self._attn_ws = ops.tree_attn_workspace(q.shape[0], q.shape[1], q.shape[2], self.device)
            self._attn_ws_rows = q.shape[0]
        return ops.tree_attn(q, self.k_cache[layer_idx], self.v_cache[layer_idx], self.kv_offset,
                             ops.make_mask(ops.UMB_MASK_DENSE, dense_mask=mask), "NHD",
                             soft_cap=float(logits_soft_cap), workspace=self._attn_ws)

    def clear(self):
        self.k_cache.zero_()
        self.v_cache.zero_()
        self.kv_offset = 0

    def set_kv_len(self, kv_len: int):
        self.kv_offset = kv_len

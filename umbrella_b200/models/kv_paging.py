"""Host-resident KV pages (BASELINE config 4): the K/V planes of the upper ``kv_host_layers`` decoder
layers live in pinned host memory and are streamed through two device staging planes, layer by layer, with
``cudaMemcpyAsync`` on a side stream ordered by events -- the weight pager's scheme (models/offload.py) applied to the
cache.  The reference keeps the whole cache on the device (umbrella/attn/cache.py:18-34: ``[L, S, Hkv, D]`` per K / V,
10.7 GB for the 70B model at 32k context); paging it out is a capability of this engine for contexts (or batches of
replicas) that do not fit, and it is PCIe-bound by construction: every forward moves the committed prefix of every paged
layer host -> device once (4 KB per position per layer for the 70B shape) and the new rows device -> host.

Per forward (layer-major, like the weight pager -- a prefill of P rows moves each paged layer's prefix once, not once per
64-row chunk):
    side stream:   H2D prefix(l) -> plane b   ...  [after the layer's kernels]  D2H new rows(l)  ->  H2D prefix(l + 2)
    main stream:   wait ready[b] -> layer l kernels (append + attention read plane b) -> record done[b]
The copy of layer l + 1 runs under layer l's kernels.  The accept-path compaction of a paged layer is a handful of row
moves in host memory (the path comes back with the 4-int result the engine reads anyway).
"""
from __future__ import annotations

import torch

from .. import ops
from ..attn.cache import KV_Cache


class PagedKV_Cache(KV_Cache):
    """KV_Cache whose planes [L_res, L) live in pinned host memory; the device tensor holds the L_res resident planes
    followed by the staging planes."""

    def __init__(self, config, max_length, device, dtype, num_layers, kv_host_layers: int, num_staging: int = 2):
        self.total_layers = num_layers
        self.host_layers = max(0, min(int(kv_host_layers), num_layers))
        self.resident_layers = num_layers - self.host_layers
        self.num_staging = num_staging if self.host_layers else 0
        super().__init__(config, max_length=max_length, device=device, dtype=dtype,
                         num_layers=self.resident_layers + self.num_staging)
        hkv, d = config.num_key_value_heads, self.head_dim
        self.k_host = torch.zeros((self.host_layers, max_length, hkv, d), dtype=dtype).pin_memory()
        self.v_host = torch.zeros((self.host_layers, max_length, hkv, d), dtype=dtype).pin_memory()
        self._path_host = torch.zeros(4 + 1024, dtype=torch.int32).pin_memory()
        self._path_evt = torch.cuda.Event()
        self._pending_gather = None          # offset of a compaction whose path is still in flight to the host
        self.pager = None

    def gather_from_result(self, result, offset, accept_len, offset_dev=None):
        if offset_dev is not None:
            raise RuntimeError("graph replay needs a device-resident cache: KV-paged models run eagerly")
        R = self.resident_layers
        if R:
            ops.kv_gather(self.k_cache[:R], self.v_cache[:R], result, offset, "LSHD")
        n = min(result.numel(), self._path_host.numel())
        self._path_host[:n].copy_(result[:n], non_blocking=True)
        self._path_evt.record()
        self._pending_gather = offset
        self.kv_offset = offset + accept_len

    def apply_pending_gather(self):
        """Compaction of the host planes: rows path[j] -> offset + j.  Called before the next forward touches them."""
        if self._pending_gather is None:
            return
        offset, self._pending_gather = self._pending_gather, None
        self._path_evt.synchronize()
        if self.pager is not None:
            self.pager.stream.synchronize()                 # the tree rows of the last forward are on the host
        m = int(self._path_host[0])
        if m and self.host_layers:
            idx = self._path_host[4:4 + m].to(torch.int64)
            self.k_host[:, offset:offset + m] = self.k_host[:, idx]
            self.v_host[:, offset:offset + m] = self.v_host[:, idx]

    def clear(self):
        super().clear()
        self._pending_gather = None


class KVPager:
    """Streams the host planes of one model through the staging planes of its PagedKV_Cache."""

    def __init__(self, cache: PagedKV_Cache):
        self.c = cache
        self.stream = torch.cuda.Stream(device=cache.device)
        self.ready = [torch.cuda.Event() for _ in range(cache.num_staging)]
        self.done = [torch.cuda.Event() for _ in range(cache.num_staging)]
        self.plane_bytes = cache.k_cache.stride(0) * cache.k_cache.element_size()
        self.stats = {"h2d_bytes": 0, "d2h_bytes": 0, "copies": 0, "forwards": 0}
        cache.pager = self

    def _plane(self, l: int) -> int:
        return self.c.resident_layers + (l - self.c.resident_layers) % self.c.num_staging

    def base_ptrs(self, l: int):
        """k / v base pointers for the C model struct such that ``base + l * plane`` is layer l's plane."""
        c = self.c
        if l < c.resident_layers or l >= c.total_layers:
            return c.k_cache.data_ptr(), c.v_cache.data_ptr()
        shift = (self._plane(l) - l) * self.plane_bytes
        return c.k_cache.data_ptr() + shift, c.v_cache.data_ptr() + shift

    def _load(self, l: int, kv_offset: int):
        c, b = self.c, (l - self.c.resident_layers) % self.c.num_staging
        with torch.cuda.stream(self.stream):
            if kv_offset:
                p, h = self._plane(l), l - c.resident_layers
                c.k_cache[p, :kv_offset].copy_(c.k_host[h, :kv_offset], non_blocking=True)
                c.v_cache[p, :kv_offset].copy_(c.v_host[h, :kv_offset], non_blocking=True)
                self.stats["h2d_bytes"] += 2 * kv_offset * c.k_cache.stride(1) * c.k_cache.element_size()
                self.stats["copies"] += 2
            self.ready[b].record(self.stream)

    def begin_forward(self, kv_offset: int, cur):
        c = self.c
        c.apply_pending_gather()
        self.stats["forwards"] += 1
        self.stream.wait_stream(cur)                        # staging planes: last readers are on the main stream
        for l in range(c.resident_layers, min(c.total_layers, c.resident_layers + c.num_staging)):
            self._load(l, kv_offset)

    def before_layer(self, l: int, kv_offset: int, cur):
        if l >= self.c.resident_layers:
            cur.wait_event(self.ready[(l - self.c.resident_layers) % self.c.num_staging])

    def after_layer(self, l: int, kv_offset: int, rows: int, cur):
        c = self.c
        if l < c.resident_layers:
            return
        b, p, h = (l - c.resident_layers) % c.num_staging, self._plane(l), l - c.resident_layers
        self.done[b].record(cur)
        self.stream.wait_event(self.done[b])
        with torch.cuda.stream(self.stream):
            c.k_host[h, kv_offset:kv_offset + rows].copy_(c.k_cache[p, kv_offset:kv_offset + rows], non_blocking=True)
            c.v_host[h, kv_offset:kv_offset + rows].copy_(c.v_cache[p, kv_offset:kv_offset + rows], non_blocking=True)
        self.stats["d2h_bytes"] += 2 * rows * c.k_cache.stride(1) * c.k_cache.element_size()
        self.stats["copies"] += 2
        if l + c.num_staging < c.total_layers:
            self._load(l + c.num_staging, kv_offset)        # next user of this plane, queued behind the D2H

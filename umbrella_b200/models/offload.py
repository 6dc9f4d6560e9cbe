"""Weight paging: host-resident layers streamed through device staging slabs.

Host-side mirror of the reference's ``LlamaOffload`` / ``LlamaAwqOffload`` (umbrella/models/llama.py:
145-219, 324-409; umbrella/models/llama_layer.py:220-258; umbrella/quantization/awq_utils.py:24-26,
54-60) with the paging engine redesigned:

  * one **pinned host slab per layer** (q|k|v, o, gate|up, down streams + the two norm vectors,
    256-byte aligned inside the slab) instead of 23 pinned tensors -> ONE cudaMemcpyAsync per layer
    instead of 23 ``copy_`` calls, the small ones latency-bound (awq_utils.py:54-60);
  * ``num_staging_layers`` (default 2, as the reference's double buffer, llama.py:192-194) device slabs;
    layer j of the offloaded range always lives in slab ``j % num_staging_layers``, so the C layer table
    is static;
  * copies run on a side stream ordered by **events** (slab free -> copy -> slab ready), never a
    device-wide ``torch.cuda.synchronize()`` (the reference has one per layer, llama.py:212);
  * the copy of layer ``j + num_staging_layers`` is enqueued as soon as layer j's kernels are, and it wraps
    around: when a forward ends, the first layers of the NEXT forward are already in flight, overlapped
    with the lm_head and the draft model's autoregressive steps;
  * layers below ``num_cache_layers`` stay resident and are used in place (the reference copies them
    into the staging buffer every time as well, llama_layer.py:251-258).

Compute is the same native path as the resident model: ``umb_llama_forward_range`` runs one layer of
<= 128 rows per call; larger inputs keep one hidden-state buffer per 128-row chunk so the weights are
streamed once per forward, not once per chunk.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib, ops
from .llama import MAX_ROWS, Llama, LlamaAwq

_ALIGN = 256
CHUNK_ROWS = 128     # rows per chunk of a paged forward (matches kChunkRows of csrc/forward.cu)
_PARTS = ("wqkv", "wo", "wgu", "wdown", "input_norm", "post_norm")


class _Paged:
    """Mixin over Llama / LlamaAwq: layers >= num_cache_layers live in pinned host memory; optionally the K/V planes of
    the upper ``kv_host_layers`` layers too (models/kv_paging.py)."""
    supports_graph = False

    def _make_kv_cache(self, **kwargs):
        host = int(kwargs.get("kv_host_layers", 0))
        if host <= 0:
            self.kv_pager = None
            return super()._make_kv_cache()
        from .kv_paging import KVPager, PagedKV_Cache
        cache = PagedKV_Cache(self.config, self.max_length, self.device, self.dtype, self.num_layers, host,
                              int(kwargs.get("kv_staging_layers", 2)))
        self.kv_pager = KVPager(cache)
        self.kv_host_layers = cache.host_layers
        return cache

    def kv_paging_stats(self) -> dict:
        p = self.kv_pager
        if p is None:
            return {}
        f = max(p.stats["forwards"], 1)
        return {"kv_host_layers": self.kv_host_layers, "h2d_bytes_per_forward": p.stats["h2d_bytes"] / f,
                "d2h_bytes_per_forward": p.stats["d2h_bytes"] / f, "copies_per_forward": p.stats["copies"] / f,
                "forwards": p.stats["forwards"]}

    def alloc(self, **kwargs):
        self.num_cache_layers = int(kwargs.get("num_cache_layers", 0))
        self.num_staging_layers = int(kwargs.get("num_staging_layers", 2))
        assert self.num_staging_layers >= 1
        super().alloc(**kwargs)
        self.num_cache_layers = max(0, min(self.num_cache_layers, self.num_layers))
        self.load_stream = torch.cuda.Stream(device=self.device)
        self.host_slabs: list[torch.Tensor | None] = [None] * self.num_layers
        self._slab_layout = None
        self._stage: list[torch.Tensor] = []
        self._stage_holds: list[int] = []
        self._ready: list[torch.cuda.Event] = []
        self._done: list[torch.cuda.Event] = []
        self._chunk_hidden = None
        self._chunk_structs = []
        self.copy_stats = {"copies": 0, "bytes": 0}
        return self

    # ------------------------------------------------------------------ weights
    def _layout(self, tensors: dict) -> dict:
        off, lay = 0, {}
        for name in _PARTS:
            nbytes = tensors[name].numel() * tensors[name].element_size()
            lay[name] = (off, nbytes, tensors[name].dtype, tuple(tensors[name].shape))
            off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        lay["bytes"] = off
        return lay

    def _set_layer(self, idx, wqkv, wo, wgu, wdown, in_norm, post_norm):
        if idx < self.num_cache_layers:
            return super()._set_layer(idx, wqkv, wo, wgu, wdown, in_norm, post_norm)
        t = dict(wqkv=wqkv, wo=wo, wgu=wgu, wdown=wdown, input_norm=in_norm.contiguous(),
                 post_norm=post_norm.contiguous())
        if self._slab_layout is None:
            self._slab_layout = self._layout(t)
            n_off = self.num_layers - self.num_cache_layers
            for _ in range(min(self.num_staging_layers, n_off)):
                self._stage.append(torch.empty(self._slab_layout["bytes"], dtype=torch.uint8, device=self.device))
                self._stage_holds.append(-1)
                self._ready.append(torch.cuda.Event())
                self._done.append(torch.cuda.Event())
        lay = self._slab_layout
        slab = torch.empty(lay["bytes"], dtype=torch.uint8).pin_memory()      # awq_utils.py:24-26 pins per tensor
        for name in _PARTS:
            off, nbytes, dtype, shape = lay[name]
            src = t[name]
            assert src.numel() * src.element_size() == nbytes and src.dtype == dtype, "layers must share one layout"
            slab[off:off + nbytes].copy_(src.reshape(-1).view(torch.uint8))
        self.host_slabs[idx] = slab
        b = (idx - self.num_cache_layers) % len(self._stage)
        base = self._stage[b].data_ptr()
        cl = self._c_layers[idx]
        cl.wqkv, cl.wo, cl.wgu, cl.wdown = (base + lay[n][0] for n in ("wqkv", "wo", "wgu", "wdown"))
        cl.input_norm, cl.post_norm = base + lay["input_norm"][0], base + lay["post_norm"][0]
        self.layer_weights[idx] = {"host_slab": slab}
        self._c_model = None

    def weight_bytes(self) -> int:
        n = self.lm_head.numel() * 2 + self.norm_weight.numel() * 2
        for i, lw in enumerate(self.layer_weights):
            n += lw["host_slab"].numel() if "host_slab" in lw else \
                sum(t.numel() * t.element_size() for t in lw.values())
        return n

    def paged_bytes_per_forward(self) -> int:
        """Host -> device bytes one forward moves (every offloaded layer's slab once)."""
        return sum(lw["host_slab"].numel() for lw in self.layer_weights if lw is not None and "host_slab" in lw)

    # ------------------------------------------------------------------ paging
    def _enqueue_load(self, b: int, j: int):
        """Copy offloaded layer j into staging slab b on the side stream, after the slab's last reader."""
        ls = self.load_stream
        ls.wait_event(self._done[b])
        with torch.cuda.stream(ls):
            self._stage[b].copy_(self.host_slabs[self.num_cache_layers + j], non_blocking=True)
            self._ready[b].record(ls)
        self._stage_holds[b] = j
        self.copy_stats["copies"] += 1
        self.copy_stats["bytes"] += self._stage[b].numel()

    def prefetch(self):
        """Start the copies of the first staged layers (normally already in flight: they are enqueued
        when the previous forward passes the same slab)."""
        n_off = self.num_layers - self.num_cache_layers
        for b in range(len(self._stage)):
            if b < n_off and self._stage_holds[b] != b:
                self._enqueue_load(b, b)

    # ------------------------------------------------------------------ forward
    def _chunk_models(self, n_chunks: int):
        """One model struct per 128-row chunk, differing only in the hidden-state buffer."""
        H = self.config.hidden_size
        if self._chunk_hidden is None or self._chunk_hidden.shape[0] < n_chunks:
            self._chunk_hidden = torch.zeros(n_chunks, MAX_ROWS, H, dtype=torch.float16, device=self.device)
            self._chunk_structs = []
        base = self._model_struct()
        while len(self._chunk_structs) < n_chunks:
            m = _lib.LlamaModel()
            C.memmove(C.byref(m), C.byref(base), C.sizeof(_lib.LlamaModel))
            m.hidden = self._chunk_hidden[len(self._chunk_structs)].data_ptr()
            self._chunk_structs.append(m)
        return self._chunk_structs[:n_chunks]

    def _model_struct(self):
        fresh = self._c_model is None
        m = super()._model_struct()
        if fresh:
            self._chunk_structs = []
        return m

    def _forward(self, input_ids, position_ids, mask: _lib.AttnMask, logits_from: int = 0, want_logits: bool = True,
                 argmax_keys=None):
        ids, pos = input_ids.reshape(-1), position_ids.reshape(-1)
        if not ids.is_cuda:
            raise RuntimeError("input_ids must be a CUDA tensor (there is no CPU path)")
        assert ids.dtype == torch.int64 and pos.dtype == torch.int64 and ids.is_contiguous() and pos.is_contiguous()
        rows, kv_offset = ids.numel(), self.kv_cache.kv_offset
        if kv_offset + rows > self.max_length:
            raise RuntimeError(f"KV cache overflow: {kv_offset} + {rows} > {self.max_length}")
        lib, cur = _lib.load(), torch.cuda.current_stream()
        st = C.c_void_p(cur.cuda_stream)
        n_chunks = (rows + CHUNK_ROWS - 1) // CHUNK_ROWS
        chunk = (rows + n_chunks - 1) // n_chunks
        spans = [(r0, min(chunk, rows - r0)) for r0 in range(0, rows, chunk)]
        models = self._chunk_models(len(spans))
        masks = []
        for r0, n in spans:
            mk = _lib.AttnMask()
            C.memmove(C.byref(mk), C.byref(mask), C.sizeof(_lib.AttnMask))
            if mk.mask_mode == ops.UMB_MASK_DENSE:
                mk.dense_mask = mask.dense_mask + r0 * mask.mask_stride
            elif mk.mask_mode == ops.UMB_MASK_CAUSAL:
                mk.prefix_len = mask.prefix_len + r0
            else:
                mk.tree_bits = mask.tree_bits + r0 * mask.tree_words * 4
            masks.append(mk)
        ip, pp = ids.data_ptr(), pos.data_ptr()
        ncache, n_stage = self.num_cache_layers, len(self._stage)
        n_off = self.num_layers - ncache
        self.prefetch()
        kvp = getattr(self, "kv_pager", None)          # host-resident KV pages (models/kv_paging.py)
        if kvp is not None:
            kvp.begin_forward(kv_offset, cur)

        def run(ci, l0, l1, embed, lg=None, lg_from=0, keys=None):
            r0, n = spans[ci]
            if kvp is not None:
                models[ci].k_cache, models[ci].v_cache = kvp.base_ptrs(l0)
            _lib.check(lib.umb_llama_forward_range(
                C.byref(models[ci]), C.c_void_p(ip + 8 * r0), C.c_void_p(pp + 8 * r0), n, kv_offset + r0,
                C.byref(masks[ci]), l0, l1, int(embed), lg, lg_from, keys, st), "llama_forward_range")

        for l in range(self.num_layers):
            j = l - ncache
            if j >= 0:
                b = j % n_stage
                if self._stage_holds[b] != j:          # only when (layers - cached) % staging != 0
                    self._enqueue_load(b, j)
                cur.wait_event(self._ready[b])
            if kvp is not None:
                kvp.before_layer(l, kv_offset, cur)
            for ci in range(len(spans)):
                run(ci, l, l + 1, l == 0)
            if kvp is not None:
                kvp.after_layer(l, kv_offset, rows, cur)
            if j >= 0 and n_off > n_stage:
                self._done[b].record(cur)
                self._enqueue_load(b, (j + n_stage) % n_off)
        logits = None
        if want_logits:
            logits = torch.empty((rows - logits_from, self.config.vocab_size), dtype=torch.float32, device=self.device)
            L = self.num_layers
            for ci, (r0, n) in enumerate(spans):
                if logits_from >= r0 + n:
                    continue
                frm = max(logits_from - r0, 0)
                out_row = r0 + frm - logits_from
                run(ci, L, L, self.num_layers == 0,
                    C.c_void_p(logits.data_ptr() + 4 * out_row * self.config.vocab_size), frm,
                    C.c_void_p(argmax_keys.data_ptr() + 8 * out_row) if argmax_keys is not None else None)
        self.kv_cache.kv_offset = kv_offset + rows
        return logits

    def forward_dyn(self, *a, **k):
        raise RuntimeError("graph replay needs resident weights: paged models run eagerly (event-ordered copies)")


class LlamaOffload(_Paged, Llama):
    """fp16 weights paged from pinned host memory (reference ``LlamaOffload``, llama.py:145-219)."""


class LlamaAwqOffload(_Paged, LlamaAwq):
    """AWQ INT4 weights paged from pinned host memory (reference ``LlamaAwqOffload``, llama.py:324-409)."""

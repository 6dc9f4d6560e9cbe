"""AutoModelLM: name -> model class (mirror of reference umbrella/models/auto_model.py:9-182 for the
Llama entries; the Qwen/Mistral/Gemma tables are out of scope).

There are no checkpoints (and no network) in this environment, so ``alloc()`` of a model built by
name attaches *synthetic* weights of that architecture unless a loader has attached real tensors
(``load_head`` / ``load_awq_layer`` / ``load_fp16_layer``) -- SURVEY.md 8d.
"""
from __future__ import annotations

import torch

from . import llama as L


class _NamedLlama:
    """Adds the reference constructor signature ``(model_name, batch_size, max_length, device, dtype)``
    and synthetic-weight allocation to the shape-driven classes of models/llama.py."""

    def __new__(cls, *a, **k):
        raise TypeError("factory only")


def _make(model_name, batch_size=1, max_length=256, device="cuda:0", dtype=torch.float16, **unused):
    entry = L.KNOWN_MODELS.get(model_name)
    if entry is None:
        raise ValueError(f"Model type '{model_name}' is not supported. "
                         f"Supported types: {[k for k, v in L.KNOWN_MODELS.items() if v]}")
    shape, awq = entry
    model = (L.LlamaAwq if awq else L.Llama)(shape, max_length=max_length, device=device, dtype=dtype,
                                             batch_size=batch_size, model_name=model_name)
    base_alloc = model.alloc

    def alloc(**kwargs):
        """Reference ``alloc(**config)`` reads num_cache_layers / exit_layer (llama.py:153,328,421)."""
        exit_layer = kwargs.get("exit_layer")
        seed = kwargs.get("synthetic_seed", 1234)
        gen_device = kwargs.get("synthetic_gen_device", str(model.device))
        if exit_layer is not None and not awq and 0 < exit_layer < model.num_layers:
            model.num_layers = exit_layer              # draft truncation (llama.py:421,450-451)
        base_alloc()
        head = L.synth_head(shape, seed, gen_device)
        model.load_head(head["embed_tokens"], head["lm_head"], head["norm"])
        for i in range(model.num_layers):
            w = L.synth_layer(shape, i, awq, seed, gen_device)
            (model.load_awq_layer if awq else model.load_fp16_layer)(i, w)
        torch.cuda.synchronize(model.device)
        return model

    model.alloc = alloc
    return model


class AutoModelLM:
    """``AutoModelLM.from_pretrained(model_name, offload=False, cuda_graph=False, **kw)``."""
    _MODEL_MAPPING = {k: v for k, v in L.KNOWN_MODELS.items() if v}

    @classmethod
    def from_pretrained(cls, model_name, offload=False, cuda_graph=False, **kwargs):
        if model_name not in cls._MODEL_MAPPING:
            raise ValueError(f"Model type '{model_name}' is not supported. "
                             f"Supported types: {list(cls._MODEL_MAPPING.keys())}")
        # offload / cuda_graph select host-paging and graph-replay variants in the reference; on a
        # 180 GB B200 every configured model is resident, and the native forward replaces per-op
        # Python, so both map to the same resident model here.
        return _make(model_name, **kwargs)

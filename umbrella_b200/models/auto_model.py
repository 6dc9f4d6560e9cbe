"""AutoModelLM: name -> model class (mirror of reference umbrella/models/auto_model.py:9-182 for the
Llama entries; the Qwen/Mistral/Gemma tables are out of scope).

``alloc()`` loads the model's safetensors checkpoint (fp16 or AutoAWQ INT4) from a local directory or the Hugging
Face cache through models/checkpoint.py -- no ``autoawq`` / ``gptqmodel`` needed -- and raises ``CheckpointNotFound``
when there is none.  Synthetic seeded weights of the architecture (there are no real weights and no network in this
environment) are strictly opt-in: ``alloc(synthetic=True)`` or a ``"synthetic": true`` /
``"synthetic_seed": N`` key in the engine config.  ``model_name`` may also be a checkpoint directory of a Llama
architecture that is not in the table (its config.json supplies the shape).
"""
from __future__ import annotations

import os

import torch

from . import llama as L


class _NamedLlama:
    """Adds the reference constructor signature ``(model_name, batch_size, max_length, device, dtype)``
    and synthetic-weight allocation to the shape-driven classes of models/llama.py."""

    def __new__(cls, *a, **k):
        raise TypeError("factory only")


def _weight_bytes(shape, awq: bool) -> int:
    per_layer = sum(k * n for k, n in L.linear_dims(shape).values())
    body = shape.num_hidden_layers * per_layer * (0.5 + 2.5 / shape.group_size if awq else 2)
    head = shape.vocab_size * shape.hidden_size * 2 * (1 if shape.tie_word_embeddings else 2)
    return int(body + head)


def _make(model_name, offload=False, batch_size=1, max_length=256, device="cuda:0", dtype=torch.float16, **unused):
    entry = L.KNOWN_MODELS.get(model_name)
    if entry is None and os.path.isfile(os.path.join(str(model_name), "config.json")):
        from .checkpoint import shape_from_config
        entry = shape_from_config(model_name)          # a local Llama-architecture checkpoint directory
    if entry is None:
        raise ValueError(f"Model type '{model_name}' is not supported. "
                         f"Supported types: {[k for k, v in L.KNOWN_MODELS.items() if v]}")
    shape, awq = entry
    if offload:
        from .offload import LlamaAwqOffload, LlamaOffload
        cls = LlamaAwqOffload if awq else LlamaOffload
    else:
        cls = L.LlamaAwq if awq else L.Llama
    model = cls(shape, max_length=max_length, device=device, dtype=dtype, batch_size=batch_size, model_name=model_name)
    base_alloc = model.alloc

    def alloc(**kwargs):
        """Reference ``alloc(**config)`` reads num_cache_layers / exit_layer (llama.py:153,328,421)."""
        exit_layer = kwargs.get("exit_layer")
        synthetic = bool(kwargs.get("synthetic", False)) or "synthetic_seed" in kwargs
        seed = kwargs.get("synthetic_seed", 1234)
        gen_device = kwargs.get("synthetic_gen_device", str(model.device))
        ckpt = None
        if not synthetic:
            from .checkpoint import CheckpointNotFound, resolve_checkpoint
            try:
                ckpt = resolve_checkpoint(model_name)
            except CheckpointNotFound as e:
                raise CheckpointNotFound(f"{e}  Pass synthetic=True (engine config: \"synthetic\": true) to run "
                                         f"seeded random weights of this architecture instead.") from None
        if exit_layer is not None and not awq and 0 < exit_layer < model.num_layers:
            model.num_layers = exit_layer              # draft truncation (llama.py:421,450-451)
        if offload:
            # The reference pages whatever ``num_cache_layers`` leaves out (configs written for 12-48 GB
            # cards).  On a 180 GB B200 a model that fits stays fully resident unless ``force_offload`` asks
            # for the paged path; a model that does not fit pages exactly as configured.
            free, _ = torch.cuda.mem_get_info(model.device)
            if not kwargs.get("force_offload", False) and _weight_bytes(shape, awq) < 0.8 * free:
                kwargs = dict(kwargs, num_cache_layers=model.num_layers)
            base_alloc(**kwargs)
        else:
            base_alloc()
        if ckpt is not None:
            from .checkpoint import load_checkpoint
            load_checkpoint(model, ckpt, awq)
        else:
            head = L.synth_head(shape, seed, gen_device)
            model.load_head(head["embed_tokens"], head["lm_head"], head["norm"])
            for i in range(model.num_layers):
                w = L.synth_layer(shape, i, awq, seed, gen_device)
                (model.load_awq_layer if awq else model.load_fp16_layer)(i, w)
        model.weights_source = ckpt or f"synthetic(seed={seed})"
        torch.cuda.synchronize(model.device)
        return model

    model.alloc = alloc
    return model


class AutoModelLM:
    """``AutoModelLM.from_pretrained(model_name, offload=False, cuda_graph=False, **kw)``."""
    _MODEL_MAPPING = {k: v for k, v in L.KNOWN_MODELS.items() if v}

    @classmethod
    def from_pretrained(cls, model_name, offload=False, cuda_graph=False, **kwargs):
        if model_name not in cls._MODEL_MAPPING and not os.path.isfile(os.path.join(str(model_name), "config.json")):
            raise ValueError(f"Model type '{model_name}' is not supported. "
                             f"Supported types: {list(cls._MODEL_MAPPING.keys())}")
        # offload=True -> the paged classes of models/offload.py (LlamaOffload / LlamaAwqOffload);
        # cuda_graph selects a graph-replay variant in the reference (static KV cache written at storage_ids) -- here
        # the engines capture the whole iteration themselves over the same native forward; the flag only selects the
        # storage-id semantics of ``graph_inference``.
        model = _make(model_name, offload=offload, **kwargs)
        model.storage_id_semantics = bool(cuda_graph)      # graph_inference writes rows where storage_ids says
        return model

"""Checkpoint loader: Hugging Face safetensors (fp16 or AutoAWQ GEMM INT4) -> the native model.

The reference loads through ``AutoModelForCausalLM.from_pretrained`` and copies tensors out of the HF modules
(umbrella/models/llama.py:39-73, 228-258; umbrella/models/llama_layer.py:27-40, 176-218;
umbrella/quantization/awq_utils.py:18-30 reads ``qweight / qzeros / scales / bias / w_bit / group_size`` off
``WQLinear_GEMM``).  That needs ``autoawq`` (and, with transformers 5.x, ``gptqmodel`` + ``accelerate``) just to
deserialise tensors the kernels repack anyway.  Here the safetensors shards are read directly:

    model.embed_tokens.weight, lm_head.weight (absent when tied), model.norm.weight
    model.layers.{i}.input_layernorm.weight, model.layers.{i}.post_attention_layernorm.weight
    model.layers.{i}.self_attn.{q,k,v,o}_proj.{weight | qweight, qzeros, scales}
    model.layers.{i}.mlp.{gate,up,down}_proj.{weight | qweight, qzeros, scales}

one decoder layer at a time (host memory stays at one layer), and handed to the existing seams
``load_head`` / ``load_awq_layer`` (-> umb_awq_repack) / ``load_fp16_layer``.

``resolve_checkpoint`` finds the files offline: a local directory, or the Hugging Face cache
(``huggingface_hub.snapshot_download(local_files_only=True)``); there is no network in this environment.
"""
from __future__ import annotations

import json
import os

import torch

from .llama import LINEAR_NAMES, LlamaShape

# our layer-dict key -> HF module path inside ``model.layers.{i}.``
HF_LINEAR = {"wq": "self_attn.q_proj", "wk": "self_attn.k_proj", "wv": "self_attn.v_proj", "wo": "self_attn.o_proj",
             "gate_proj": "mlp.gate_proj", "up_proj": "mlp.up_proj", "down_proj": "mlp.down_proj"}


class CheckpointNotFound(FileNotFoundError):
    pass


def resolve_checkpoint(model_name: str) -> str:
    """Directory holding config.json + *.safetensors for ``model_name`` (a path or a hub id), offline."""
    if os.path.isdir(model_name):
        return model_name
    env = os.environ.get("UMBRELLA_CHECKPOINT_DIR")
    if env:
        cand = os.path.join(env, model_name.replace("/", "--"))
        if os.path.isdir(cand):
            return cand
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(model_name, local_files_only=True, allow_patterns=["*.json", "*.safetensors"])
    except Exception as e:
        raise CheckpointNotFound(
            f"no local checkpoint for '{model_name}' (looked at the path itself, $UMBRELLA_CHECKPOINT_DIR and the "
            f"Hugging Face cache; this environment is offline): {e!r}") from None


def shape_from_config(path: str) -> tuple[LlamaShape, bool]:
    """config.json -> (LlamaShape, is_awq).  Mirrors what the reference reads off LlamaConfig (llama.py:24-33)
    plus ``quantization_config`` (AutoAWQ: quant_method awq, bits 4, group_size 128, version gemm)."""
    cfg = json.load(open(os.path.join(path, "config.json")))
    rope_params = cfg.get("rope_parameters") or {}
    rope_scaling = cfg.get("rope_scaling")
    if rope_scaling is None and rope_params.get("rope_type", "default") != "default":
        rope_scaling = {k: v for k, v in rope_params.items() if k != "rope_theta"}
    eos = cfg.get("eos_token_id", (128001, 128008, 128009))
    eos = list(eos) if isinstance(eos, (list, tuple)) else [eos]
    q = cfg.get("quantization_config") or {}
    awq = q.get("quant_method") == "awq"
    if awq:
        if q.get("bits", 4) != 4 or q.get("group_size", 128) != 128 or str(q.get("version", "gemm")).lower() != "gemm":
            raise ValueError(f"unsupported AWQ flavour {q}: the kernels implement 4-bit, group 128, GEMM layout")
    shape = LlamaShape(hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                       num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
                       num_key_value_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]),
                       vocab_size=cfg["vocab_size"], rms_norm_eps=cfg.get("rms_norm_eps", 1e-5),
                       rope_theta=cfg.get("rope_theta", rope_params.get("rope_theta", 10000.0)),
                       max_position_embeddings=cfg.get("max_position_embeddings", 8192), rope_scaling=rope_scaling,
                       tie_word_embeddings=bool(cfg.get("tie_word_embeddings", False)),
                       group_size=q.get("group_size", 128), eos_token_id=eos)
    return shape, awq


class SafetensorsIndex:
    """tensor name -> shard file, over ``model.safetensors.index.json`` or a single ``model.safetensors``."""

    def __init__(self, path: str):
        self.path = path
        idx = os.path.join(path, "model.safetensors.index.json")
        if os.path.exists(idx):
            self.where = dict(json.load(open(idx))["weight_map"])
        else:
            from safetensors import safe_open
            files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
            if not files:
                raise CheckpointNotFound(f"no *.safetensors under {path}")
            self.where = {}
            for f in files:
                with safe_open(os.path.join(path, f), framework="pt") as h:
                    for k in h.keys():
                        self.where[k] = f
        self._open = {}

    def __contains__(self, name: str) -> bool:
        return name in self.where

    def get(self, name: str) -> torch.Tensor:
        from safetensors import safe_open
        f = self.where[name]
        h = self._open.get(f)
        if h is None:
            h = self._open[f] = safe_open(os.path.join(self.path, f), framework="pt")
        return h.get_tensor(name)

    def close(self):
        self._open.clear()


def read_layer(index: SafetensorsIndex, i: int, awq: bool) -> dict:
    """The tensors of decoder layer ``i`` in the form load_awq_layer / load_fp16_layer take."""
    pre = f"model.layers.{i}."
    out = {}
    for ours, hf in HF_LINEAR.items():
        if awq:
            qw, qz, sc = (index.get(pre + hf + s) for s in (".qweight", ".qzeros", ".scales"))
            if (pre + hf + ".bias") in index:
                raise ValueError(f"{pre + hf}: AWQ linears with a bias are not supported by the fused streams")
            if qw.dtype != torch.int32 or qz.dtype != torch.int32:
                raise ValueError(f"{pre + hf}: qweight / qzeros must be int32 (AutoAWQ GEMM layout)")
            out[ours] = (qw.contiguous(), qz.contiguous(), sc.to(torch.float16).contiguous())
        else:
            out[ours] = index.get(pre + hf + ".weight").to(torch.float16)
    out["input_layernorm"] = index.get(pre + "input_layernorm.weight").to(torch.float16)
    out["post_attention_layernorm"] = index.get(pre + "post_attention_layernorm.weight").to(torch.float16)
    return out


def read_head(index: SafetensorsIndex, tied: bool) -> dict:
    embed = index.get("model.embed_tokens.weight").to(torch.float16)
    lm_head = None if (tied or "lm_head.weight" not in index) else index.get("lm_head.weight").to(torch.float16)
    return {"embed_tokens": embed, "lm_head": lm_head, "norm": index.get("model.norm.weight").to(torch.float16)}


def check_layer_shapes(shape: LlamaShape, w: dict, awq: bool, layer: int):
    from .llama import linear_dims
    for name, (k, n) in linear_dims(shape).items():
        if awq:
            qw, qz, sc = w[name]
            want = ((k, n // 8), (k // shape.group_size, n // 8), (k // shape.group_size, n))
            got = (tuple(qw.shape), tuple(qz.shape), tuple(sc.shape))
        else:
            want, got = ((n, k),), (tuple(w[name].shape),)
        if want != got:
            raise ValueError(f"layer {layer} {name}: checkpoint shapes {got} do not match the config {want}")


def load_checkpoint(model, path: str, awq: bool, num_layers: int | None = None):
    """Fill an allocated model (models/llama.py classes, resident or paged) from the safetensors under ``path``."""
    index = SafetensorsIndex(path)
    head = read_head(index, model.config.tie_word_embeddings)
    model.load_head(head["embed_tokens"], head["lm_head"], head["norm"])
    n = model.num_layers if num_layers is None else num_layers
    for i in range(n):
        w = read_layer(index, i, awq)
        check_layer_shapes(model.config, w, awq, i)
        (model.load_awq_layer if awq else model.load_fp16_layer)(i, w)
        del w
    index.close()
    return model


def write_synthetic_checkpoint(path: str, shape: LlamaShape, awq: bool, seed: int = 1234, shard_layers: int = 2):
    """Write a seeded synthetic model in the on-disk format above (tests, examples: there are no real weights
    offline).  Layers go into shards of ``shard_layers`` layers with a model.safetensors.index.json."""
    from safetensors.torch import save_file
    from .llama import synth_head, synth_layer
    os.makedirs(path, exist_ok=True)
    cfg = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": shape.hidden_size,
           "intermediate_size": shape.intermediate_size, "num_hidden_layers": shape.num_hidden_layers,
           "num_attention_heads": shape.num_attention_heads, "num_key_value_heads": shape.num_key_value_heads,
           "vocab_size": shape.vocab_size, "rms_norm_eps": shape.rms_norm_eps, "rope_theta": shape.rope_theta,
           "max_position_embeddings": shape.max_position_embeddings, "rope_scaling": shape.rope_scaling,
           "tie_word_embeddings": shape.tie_word_embeddings, "eos_token_id": list(shape.eos_token_id),
           "torch_dtype": "float16"}
    if awq:
        cfg["quantization_config"] = {"quant_method": "awq", "bits": 4, "group_size": shape.group_size, "version": "gemm",
                                      "zero_point": True}
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"), indent=1)
    weight_map = {}

    def flush(tensors, fname):
        save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(path, fname))
        weight_map.update({k: fname for k in tensors})

    head = synth_head(shape, seed)
    t = {"model.embed_tokens.weight": head["embed_tokens"], "model.norm.weight": head["norm"]}
    if head["lm_head"] is not None:
        t["lm_head.weight"] = head["lm_head"]
    flush(t, "model-head.safetensors")
    t = {}
    for i in range(shape.num_hidden_layers):
        w = synth_layer(shape, i, awq, seed)
        pre = f"model.layers.{i}."
        for ours, hf in HF_LINEAR.items():
            if awq:
                t[pre + hf + ".qweight"], t[pre + hf + ".qzeros"], t[pre + hf + ".scales"] = w[ours]
            else:
                t[pre + hf + ".weight"] = w[ours]
        t[pre + "input_layernorm.weight"] = w["input_layernorm"]
        t[pre + "post_attention_layernorm.weight"] = w["post_attention_layernorm"]
        if (i + 1) % shard_layers == 0 or i + 1 == shape.num_hidden_layers:
            flush(t, f"model-{i // shard_layers:05d}.safetensors")
            t = {}
    json.dump({"metadata": {}, "weight_map": weight_map}, open(os.path.join(path, "model.safetensors.index.json"), "w"))
    return path


assert set(HF_LINEAR) == set(LINEAR_NAMES)

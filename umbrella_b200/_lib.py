"""ctypes binding of libumbrella_b200.so (the C ABI declared in include/umbrella_b200.h).

The library is the product: there is no PyTorch / CPU fallback behind these calls.  Loading fails
loudly when the shared object is missing (build it with ``python -m umbrella_b200.build`` or
``__graft_entry__.build()``), and every entry point raises ``RuntimeError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# UMB_LIB_PATH selects an experiment build of the same library (umbrella_b200/build.py --variant)
LIB_PATH = os.environ.get("UMB_LIB_PATH") or os.path.join(HERE, "libumbrella_b200.so")

c_void_p, c_int, c_int64, c_size_t, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float

UMB_EPI_STORE, UMB_EPI_RESIDUAL, UMB_EPI_SILU_MUL, UMB_EPI_LOGITS, UMB_EPI_ROPE_QKV = 0, 1, 2, 3, 4
UMB_MASK_TREE, UMB_MASK_DENSE, UMB_MASK_CAUSAL = 0, 1, 2
UMB_SAMPLE_STATIC, UMB_SAMPLE_DYNAMIC, UMB_SAMPLE_ARGMAX = 0, 1, 2


class Epilogue(C.Structure):
    _fields_ = [("mode", c_int), ("bias", c_void_p), ("residual", c_void_p), ("ld_residual", c_int64),
                ("argmax", c_void_p), ("rope", c_void_p)]


class RopeEpilogue(C.Structure):
    _fields_ = [("q_out", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p), ("cos", c_void_p), ("sin", c_void_p),
                ("position_ids", c_void_p), ("n_dev", c_void_p), ("row_offset", c_int), ("kv_offset", c_int),
                ("num_q_heads", c_int), ("num_kv_heads", c_int), ("head_dim", c_int), ("kv_tok_stride", c_int64),
                ("kv_head_stride", c_int64)]


class AttnMask(C.Structure):
    _fields_ = [("mask_mode", c_int), ("prefix_len", c_int), ("dyn", c_void_p), ("tree_bits", c_void_p),
                ("tree_words", c_int), ("dense_mask", c_void_p), ("mask_stride", c_int64), ("kv_len_bound", c_int)]


class LlamaLayer(C.Structure):
    _fields_ = [("input_norm", c_void_p), ("post_norm", c_void_p), ("wqkv", c_void_p), ("wo", c_void_p),
                ("wgu", c_void_p), ("wdown", c_void_p)]


class LlamaModel(C.Structure):
    _fields_ = [("hidden_size", c_int), ("intermediate_size", c_int), ("num_layers", c_int), ("num_q_heads", c_int),
                ("num_kv_heads", c_int), ("head_dim", c_int), ("vocab", c_int), ("rms_eps", c_float), ("awq", c_int),
                ("max_length", c_int), ("layers", C.POINTER(LlamaLayer)), ("embed", c_void_p),
                ("final_norm", c_void_p), ("lm_head", c_void_p), ("cos", c_void_p), ("sin", c_void_p),
                ("k_cache", c_void_p), ("v_cache", c_void_p), ("hidden", c_void_p), ("normed", c_void_p),
                ("qkv", c_void_p), ("q_rot", c_void_p), ("attn_out", c_void_p), ("act", c_void_p),
                ("gemm_ws", c_void_p), ("gemm_ws_bytes", c_size_t), ("attn_ws", c_void_p), ("attn_ws_bytes", c_size_t)]


# name -> (restype, argtypes); must list every symbol include/umbrella_b200.h declares
SIGNATURES = {
    "umb_version": (C.c_char_p, []),
    "umb_error_string": (C.c_char_p, [c_int]),
    "umb_launch_count": (C.c_ulonglong, []),
    "umb_launch_count_add": (None, [C.c_ulonglong]),
    "umb_awq_packed_bytes": (c_size_t, [c_int, c_int, c_int]),
    "umb_awq_repack": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                               c_int, c_void_p]),
    "umb_awq_dequantize_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "umb_awq_unpack_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "umb_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "umb_awq_gemm_f16": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                 C.POINTER(Epilogue), c_void_p, c_size_t, c_void_p]),
    "umb_gemm_f16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                             C.POINTER(Epilogue), c_void_p, c_size_t, c_void_p]),
    "umb_tree_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "umb_tree_attn_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64,
                                  c_int64, C.POINTER(AttnMask), c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "umb_rope_append_f16": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "umb_tree_accept": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p]),
    "umb_kv_gather": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p, c_int,
                              c_void_p]),
    "umb_rmsnorm_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "umb_silu_mul_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "umb_add_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "umb_embedding_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "umb_argmax_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "umb_argmax_keys_to_ids": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "umb_debug_gemm_trace": (c_int, [c_void_p, c_int]),
    "umb_probe_umma": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "umb_awq_gemm_f16_simt": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "umb_llama_gemm_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "umb_llama_forward": (c_int, [C.POINTER(LlamaModel), c_void_p, c_void_p, c_int, c_int, C.POINTER(AttnMask),
                                  c_void_p, c_int, c_void_p, c_void_p]),
    "umb_llama_forward_range": (c_int, [C.POINTER(LlamaModel), c_void_p, c_void_p, c_int, c_int, C.POINTER(AttnMask),
                                        c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "umb_llama_forward_dyn": (c_int, [C.POINTER(LlamaModel), c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                      c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "umb_sample_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_int, c_void_p,
                               c_int, c_void_p, c_float, c_void_p, c_void_p]),
    "umb_topk_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "umb_store_tokens_gather": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "umb_store_tokens": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "umb_dyn_tree_expand": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library (once).  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library is the only implementation of this path. "
            "Build it with `python -m umbrella_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().umb_error_string(rc).decode()
        raise RuntimeError(f"umbrella_b200 {what} failed: status {rc} ({msg})")


def launch_count() -> int:
    return int(load().umb_launch_count())

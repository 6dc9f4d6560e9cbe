"""AwqLinear: host-side mirror of reference umbrella/quantization/awq_utils.py:5-86.

Same container surface (``init_parameters / empty_like / to / copy / apply`` over
``qweight / qzeros / scales / bias``); ``apply`` routes to the tcgen05 AWQ GEMM
(umb_awq_gemm_f16) instead of ``awq_ext.gemm_forward_cuda`` / ``dequantize_weights_cuda`` + matmul.
The checkpoint triplet is repacked once into the K-major TMA stream when it reaches the device
(``to`` / ``copy``), the seam SURVEY.md 8b names.
"""
from __future__ import annotations

import torch

from .. import ops


_TRIPLET = ("qweight", "qzeros", "scales")          # the AutoAWQ GEMM-layout checkpoint tensors of one linear


class AwqLinear:
    def __init__(self):
        self.in_features = self.out_features = self.w_bit = self.group_size = 0
        for name in _TRIPLET + ("bias",):
            setattr(self, name, None)
        self.packed: torch.Tensor = None       # device-resident repacked stream
        self._workspace = None

    def _meta(self, module):
        for name in ("in_features", "out_features", "w_bit", "group_size"):
            setattr(self, name, getattr(module, name))
        if (self.w_bit, self.group_size) != (4, 128):
            raise ValueError(f"AwqLinear: only w_bit=4, group_size=128 is implemented (got {self.w_bit}, {self.group_size})")

    def _adopt(self, module, transform):
        """Take the checkpoint tensors of ``module`` through ``transform`` (pin / zeros_like)."""
        self._meta(module)
        for name in _TRIPLET:
            setattr(self, name, transform(getattr(module, name).detach()))
        bias = getattr(module, "bias", None)
        self.bias = None if bias is None else transform(bias.detach())
        self.packed = None

    def init_parameters(self, module):
        """``module``: anything with WQLinear_GEMM's attributes (awq_utils.py:18-30); host copies are pinned so
        that ``copy`` into a device-side twin is an asynchronous DMA (awq_utils.py:24-26)."""
        self._adopt(module, (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t))

    def empty_like(self, module):
        self._adopt(module, torch.zeros_like)

    def _repack(self):
        if self.qweight.is_cuda:
            self.packed = ops.awq_repack(self.qweight, self.qzeros, self.scales, self.group_size, packed=self.packed)

    def _each(self):
        return [n for n in _TRIPLET + ("bias",) if getattr(self, n) is not None]

    def to(self, device, non_blocking=True):
        for name in self._each():
            setattr(self, name, getattr(self, name).to(device, non_blocking=non_blocking))
        self._repack()

    def copy(self, module: "AwqLinear", non_blocking=True):
        for name in self._each():
            getattr(self, name).copy_(getattr(module, name), non_blocking=non_blocking)
        self._repack()

    def apply(self, x: torch.Tensor) -> torch.Tensor:
        """x [bsz, q_len, K] fp16 -> [bsz, q_len, N] (awq_utils.py:63-86).  Rows go through the GEMM in
        chunks of 128 (the reference switches to dequantise + matmul at >= 1024 rows; here long
        prefills simply take more chunks)."""
        if self.packed is None:
            if self.qweight is None or not self.qweight.is_cuda:
                raise RuntimeError("AwqLinear.apply: weights are not on a CUDA device (there is no CPU path)")
            self._repack()
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        K, N = self.in_features, self.out_features
        out = torch.empty((x2.shape[0], N), dtype=torch.float16, device=x.device)
        if self._workspace is None:
            self._workspace = ops.gemm_workspace(128, K, N, x.device)
        for r0 in range(0, x2.shape[0], 128):
            r1 = min(x2.shape[0], r0 + 128)
            ops.awq_gemm(x2[r0:r1], self.packed, K, N, bias=self.bias, out=out[r0:r1], workspace=self._workspace,
                         group=self.group_size)
        return out.reshape(*shape[:-1], N)

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


class AwqLinear:
    def __init__(self):
        self.in_features = 0
        self.out_features = 0
        self.w_bit = 0
        self.group_size = 0
        self.qweight: torch.Tensor = None
        self.qzeros: torch.Tensor = None
        self.scales: torch.Tensor = None
        self.bias: torch.Tensor = None
        self.packed: torch.Tensor = None       # device-resident repacked stream
        self._workspace = None

    def _meta(self, module):
        self.in_features = module.in_features
        self.out_features = module.out_features
        self.w_bit = module.w_bit
        self.group_size = module.group_size
        if self.w_bit != 4 or self.group_size != 128:
            raise ValueError(f"AwqLinear: only w_bit=4, group_size=128 is implemented (got {self.w_bit}, {self.group_size})")

    def init_parameters(self, module):
        """``module``: anything with WQLinear_GEMM's attributes (awq_utils.py:18-30)."""
        self._meta(module)
        pin = (lambda t: t.detach().pin_memory()) if torch.cuda.is_available() else (lambda t: t.detach())
        self.qweight, self.qzeros, self.scales = pin(module.qweight), pin(module.qzeros), pin(module.scales)
        self.bias = module.bias.detach() if getattr(module, "bias", None) is not None else None
        self.packed = None

    def empty_like(self, module):
        self._meta(module)
        self.qweight = torch.zeros_like(module.qweight.detach())
        self.qzeros = torch.zeros_like(module.qzeros.detach())
        self.scales = torch.zeros_like(module.scales.detach())
        if getattr(module, "bias", None) is not None:
            self.bias = torch.zeros_like(module.bias.detach())
        self.packed = None

    def _repack(self):
        if self.qweight.is_cuda:
            self.packed = ops.awq_repack(self.qweight, self.qzeros, self.scales, self.group_size, packed=self.packed)

    def to(self, device, non_blocking=True):
        self.qweight = self.qweight.to(device, non_blocking=non_blocking)
        self.qzeros = self.qzeros.to(device, non_blocking=non_blocking)
        self.scales = self.scales.to(device, non_blocking=non_blocking)
        if self.bias is not None:
            self.bias = self.bias.to(device, non_blocking=non_blocking)
        self._repack()

    def copy(self, module: "AwqLinear", non_blocking=True):
        self.qweight.copy_(module.qweight, non_blocking=non_blocking)
        self.qzeros.copy_(module.qzeros, non_blocking=non_blocking)
        self.scales.copy_(module.scales, non_blocking=non_blocking)
        if self.bias is not None:
            self.bias.copy_(module.bias, non_blocking=non_blocking)
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

"""Per-format linear "methods": how a projection's weight is stored and which GEMM applies it.

Mirrors the reference's strategy registry (lite_llama/models/quantization/methods/__init__.py:21-63):
``fp16`` -> dense GEMM (unquantized.py:21-22), ``int4`` -> w4a16 (w4a16.py:15-43), ``int8`` / ``fp8``
-> w8a16 (w8a16.py:37-112), ``smoothquant`` -> w8a8 (w8a8.py:15-37).
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from ..kernels import PackedW4, linear, smoothquant_matmul, w4a16_matmul, w4a16_matmul_packed, w4a16_pack, w8a16_matmul
from ..kernels.quantization import interleave_gate_up
from .params import quantize_fp8_per_channel, quantize_int4_groupwise, quantize_int8_per_channel

LINEAR_METHODS = ("fp16", "int4", "int8", "fp8", "smoothquant")


@dataclass
class LinearWeights:
    """Storage of one ``[N, K]`` projection in one of the supported formats."""

    method: str
    out_features: int
    in_features: int
    weight: torch.Tensor | None = None      # fp16/bf16 [N, K]
    qweight: torch.Tensor | None = None     # int32 [N, K/8] | uint8/int8 [N, K]
    scales: torch.Tensor | None = None
    zeros: torch.Tensor | None = None
    bias: torch.Tensor | None = None
    group_size: int = 128
    packed: PackedW4 | None = None          # int4: kernel-native repack (replaces qweight/scales/zeros)
    swiglu: bool = False                    # gate|up projection stored row-interleaved: the GEMM applies silu(gate) * up

    def nbytes(self) -> int:
        total = 0
        for t in (self.weight, self.qweight, self.scales, self.zeros, self.bias):
            if t is not None:
                total += t.numel() * t.element_size()
        if self.packed is not None:
            total += self.packed.nbytes()
        return total


def make_linear_weights(weight: torch.Tensor, method: str, bias: torch.Tensor | None = None,
                        group_size: int = 128, fuse_swiglu: bool = False) -> LinearWeights:
    """Convert a float ``[N, K]`` weight into ``method``'s storage (the ``--quantization`` path of the
    reference: methods/w4a16.py:39-43, w8a16.py:95-112, w8a8.py:30-37)."""
    n, k = weight.shape
    if method == "fp16":
        return LinearWeights(method, n, k, weight=weight.contiguous(), bias=bias)
    if method == "int4":
        q, s, z = quantize_int4_groupwise(weight, group_size)
        if q.is_cuda and os.environ.get("B2L_W4_LAYOUT", "packed") != "reference":
            # load-time repack into the kernel-native layout (the reference's `_layout` seam); the reference-layout
            # tensors are not kept, so a projection is stored exactly once
            fuse = fuse_swiglu and n % 2 == 0 and os.environ.get("B2L_FUSE_SWIGLU", "1") != "0"
            if fuse:   # [gate | up] rows -> (gate_i, up_i) pairs: the GEMM epilogue applies the activation
                q, s, z = interleave_gate_up(q), interleave_gate_up(s), interleave_gate_up(z)
                bias = interleave_gate_up(bias) if bias is not None else None
            return LinearWeights(method, n, k, bias=bias, group_size=group_size,
                                 packed=w4a16_pack(q, s, z, group_size=group_size), swiglu=fuse)
        return LinearWeights(method, n, k, qweight=q, scales=s, zeros=z, bias=bias, group_size=group_size)
    if method in ("int8", "smoothquant"):
        q, s = quantize_int8_per_channel(weight)
        return LinearWeights(method, n, k, qweight=q, scales=s, bias=bias)
    if method == "fp8":
        q, s = quantize_fp8_per_channel(weight)
        return LinearWeights(method, n, k, qweight=q, scales=s, bias=bias)
    raise ValueError(f"unknown linear method {method!r}; expected one of {LINEAR_METHODS}")


def apply_linear(x: torch.Tensor, w: LinearWeights) -> torch.Tensor:
    """``x @ W.T (+ bias)`` through the GEMM that matches ``w.method``."""
    if w.method == "fp16":
        return linear(x, w.weight, w.bias)
    if w.method == "int4":
        if w.packed is not None:
            return w4a16_matmul_packed(x, w.packed, bias=w.bias, swiglu=w.swiglu)
        return w4a16_matmul(x, w.qweight, w.scales, w.zeros, group_size=w.group_size, bias=w.bias)
    if w.method in ("int8", "fp8"):
        return w8a16_matmul(x, w.qweight, w.scales, group_n=1, group_k=w.in_features, bias=w.bias)
    if w.method == "smoothquant":
        return smoothquant_matmul(x, w.qweight, w.scales, bias=w.bias)
    raise ValueError(f"unknown linear method {w.method!r}")

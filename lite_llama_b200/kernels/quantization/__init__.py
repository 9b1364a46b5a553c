"""Quantised GEMMs with the reference's signatures
(reference: lite_llama/kernels/quantization/__init__.py:13-17)."""

from .w4a16 import PackedW4, interleave_gate_up, w4a16_matmul, w4a16_matmul_packed, w4a16_pack
from .w8a16 import w8a16_matmul
from .w8a8 import quantize_activations_int8, smoothquant_matmul, smoothquant_matmul_prequant

__all__ = ["w8a16_matmul", "w4a16_matmul", "smoothquant_matmul", "smoothquant_matmul_prequant", "quantize_activations_int8", "w4a16_pack", "w4a16_matmul_packed", "PackedW4", "interleave_gate_up"]

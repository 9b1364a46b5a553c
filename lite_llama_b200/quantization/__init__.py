"""Weight quantisers and per-format linear methods (the reference's ``models/quantization`` seam:
lite_llama/models/quantization/params/*.py and methods/*.py). The packers define the weight layouts
the CUDA GEMMs consume bit-for-bit."""

from .params import (
    quantize_fp8_per_channel,
    quantize_int4_groupwise,
    quantize_int8_groupwise,
    quantize_int8_per_channel,
)
from .methods import LINEAR_METHODS, LinearWeights, apply_linear, make_linear_weights

__all__ = [
    "quantize_fp8_per_channel",
    "quantize_int4_groupwise",
    "quantize_int8_groupwise",
    "quantize_int8_per_channel",
    "LINEAR_METHODS",
    "LinearWeights",
    "apply_linear",
    "make_linear_weights",
]

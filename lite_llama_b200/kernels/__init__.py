"""B200 kernels behind the ``lite_llama.kernels`` names
(reference export list: lite_llama/kernels/__init__.py:22-39).

Every callable keeps the reference's signature and argument meaning; all of them launch
hand-written sm_100a CUDA from ``libb200lite.so`` on the current CUDA stream.
"""

from .flashattention2_nopad import flash_attention2_no_pad
from .flashdecoding import flash_decoding
from .fused_moe import fused_moe, moe_align_block_size, route_topk
from .linear import linear
from .rope_emb import rope_emb_forward, rope_emb_inplace, rope_kv_write
from .sampling import argmax_rows
from .skip_rmsnorm import can_fuse_int8_quant, skip_rmsnorm, skip_rmsnorm_quant_int8
from .swiglu import swiglu_forward
from .update_kv_buffer import update_kv_buffer, update_kv_buffer_split
from .update_kv_index import update_kv_index
from .quantization import PackedW4, quantize_activations_int8, smoothquant_matmul, smoothquant_matmul_prequant, w4a16_matmul, w4a16_matmul_packed, w4a16_pack, w8a16_matmul

__all__ = [
    "flash_attention2_no_pad",
    "flash_decoding",
    "fused_moe",
    "moe_align_block_size",
    "route_topk",
    "linear",
    "argmax_rows",
    "rope_emb_forward",
    "rope_emb_inplace",
    "rope_kv_write",
    "skip_rmsnorm",
    "skip_rmsnorm_quant_int8",
    "can_fuse_int8_quant",
    "swiglu_forward",
    "update_kv_buffer",
    "update_kv_buffer_split",
    "update_kv_index",
    "w4a16_matmul",
    "w8a16_matmul",
    "smoothquant_matmul",
    "smoothquant_matmul_prequant",
    "quantize_activations_int8",
    "w4a16_pack",
    "w4a16_matmul_packed",
    "PackedW4",
]

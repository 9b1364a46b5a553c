"""Runtime weight quantisers producing exactly the reference's storage layouts.

* int4 group-wise: lite_llama/models/quantization/params/int4.py:12-49
* int8 per-channel / group-wise: .../params/int8.py:12-53
* fp8-e4m3 per-channel: .../params/fp8.py:16-30
"""

from __future__ import annotations

import torch

FP8_E4M3_MAX = 448.0


def quantize_int4_groupwise(weight: torch.Tensor, group_size: int = 128):
    """``[N, K]`` float -> ``(qweight [N, K/8] int32, scales [N, K/g] fp32, zeros [N, K/g] fp32)``.

    Asymmetric 4-bit: ``scale = (max - min) / 14`` (floored at 1e-5), ``zero = round(-min / scale)``
    clamped to [0, 15], codes ``round(w / scale + zero)`` clamped to [0, 15]; eight codes per int32,
    code ``j`` of word ``i`` in bits ``[4j, 4j+4)`` = input channel ``8i + j``.
    """
    n, k = weight.shape
    if k % group_size != 0:
        raise ValueError(f"in_features {k} must be a multiple of group_size {group_size}")
    w = weight.float().reshape(n, k // group_size, group_size)
    lo, hi = w.amin(dim=-1), w.amax(dim=-1)
    scale = (hi - lo).clamp(min=1e-5) / 14.0
    zero = (-lo / scale).round().clamp(0, 15)
    codes = (w / scale.unsqueeze(-1) + zero.unsqueeze(-1)).round().clamp(0, 15).to(torch.int32)
    codes = codes.reshape(n, -1, 8)
    shifts = torch.arange(8, device=codes.device, dtype=torch.int32) * 4
    packed = (codes << shifts[None, None, :]).sum(dim=-1)
    return packed.to(torch.int32), scale.float(), zero.float()


def quantize_int8_per_channel(weight: torch.Tensor):
    """Symmetric int8, one scale per output row: ``scale = absmax / 127`` (1 where the row is 0)."""
    scale = weight.abs().amax(dim=-1, keepdim=True).float() / 127.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    q = (weight.float() / scale).round().clamp_(-127, 127).to(torch.int8)
    return q, scale


def quantize_int8_groupwise(weight: torch.Tensor, group_size: int = 128):
    """Symmetric int8 with one scale per ``group_size`` input channels."""
    k = weight.shape[-1]
    if k % group_size != 0:
        raise ValueError(f"in_features {k} must be a multiple of group_size {group_size}")
    w = weight.float().unflatten(-1, (k // group_size, group_size))
    scale = w.abs().amax(dim=-1) / 127.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    q = (w / scale.unsqueeze(-1)).round().clamp_(-127, 127).to(torch.int8)
    return q.flatten(-2), scale


def quantize_fp8_per_channel(weight: torch.Tensor):
    """fp8-e4m3 bytes in a uint8 container, one scale per output row (``absmax / 448``)."""
    scale = weight.abs().amax(dim=-1, keepdim=True).float() / FP8_E4M3_MAX
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    q = (weight.float() / scale).clamp_(-FP8_E4M3_MAX, FP8_E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), scale


def quantize_fp8_blockwise(weight: torch.Tensor, block: int = 128):
    """``[..., N, K]`` float -> (uint8 e4m3 bit patterns, fp32 scales ``[..., ceil(N/block), ceil(K/block)]``) with one
    scale per ``block x block`` tile = tile absmax / 448: the fine-grained FP8 format of the Qwen3 / DeepSeek
    checkpoints (lite_llama/models/quantization/config.py:33, executor/weight_utils.py:59-72 dequantises
    ``code * scale`` per tile)."""
    *lead, n, k = weight.shape
    nb, kb = (n + block - 1) // block, (k + block - 1) // block
    w = weight.float()
    if nb * block != n or kb * block != k:
        w = torch.nn.functional.pad(w, (0, kb * block - k, 0, nb * block - n))
    tiles = w.reshape(*lead, nb, block, kb, block)
    amax = tiles.abs().amax(dim=(-3, -1))
    scales = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    q = (tiles / scales[..., :, None, :, None]).reshape(*lead, nb * block, kb * block)[..., :n, :k]
    return q.to(torch.float8_e4m3fn).view(torch.uint8).contiguous(), scales.contiguous()


def reblock_scales(scales: torch.Tensor, block: int, new_block: int, dim: int) -> torch.Tensor:
    """Refine block scales along ``dim`` from ``block`` to ``new_block`` elements per scale (``new_block`` divides
    ``block``): every sub-block inherits its parent's scale, so ``code * scale`` is unchanged. This is what lets a
    tensor-parallel shard start in the middle of a 128-wide checkpoint block (Qwen3-30B-A3B: 768 / 4 = 192 expert
    columns per rank) -- the reference requires shards to hold whole scale blocks (models/linear.py:164-177) and
    therefore loads this model at TP = 2 only."""
    if block % new_block != 0:
        raise ValueError(f"new block {new_block} must divide {block}")
    return scales.repeat_interleave(block // new_block, dim=dim)

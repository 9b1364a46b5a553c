"""Load-time ingestion of real 4-bit checkpoints: AutoAWQ "GEMM" tensors and AutoGPTQ tensors ->
the reference's W4A16 layout (lite_llama/kernels/quantization/w4a16.py:9-12: ``qweight int32 [N, K/8]`` with nibble j
of word i = input channel 8i + j, ``scales`` / ``zeros`` fp32 ``[N, K/g]``), which ``w4a16_matmul`` consumes as is and
``w4a16_pack`` turns into the kernel-native layout.

The reference maps ``quant_method: awq | gptq`` to its int4 format (models/quantization/config.py:39-44) but ships no
reader for the checkpoints' own tensor layouts; these two functions are that reader. Both formats store the weight
input-major (``[K, N]``-shaped, eight 4-bit values per int32):

* AutoAWQ GEMM: ``qweight int32 [K, N/8]`` and ``qzeros int32 [K/g, N/8]`` pack eight OUTPUT channels per word in the
  order ``[0, 2, 4, 6, 1, 3, 5, 7]`` (nibble p of word c = channel ``8c + order[p]``); ``scales fp16 [K/g, N]``;
  ``w[k, n] = (q[k, n] - z[k // g, n]) * s[k // g, n]``.
* AutoGPTQ (no act-order): ``qweight int32 [K/8, N]`` packs eight INPUT channels per word in natural order;
  ``qzeros int32 [K/g, N/8]`` packs output channels in natural order and stores ``zero - 1``; ``scales fp16 [K/g, N]``.

Pure tensor shuffles (exact); run once per projection at load time, on whichever device the tensors live.
"""

from __future__ import annotations

import torch

AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)


def _unpack_last_dim(packed: torch.Tensor, order=tuple(range(8))) -> torch.Tensor:
    """int32 ``[..., C]`` -> uint8-valued int32 ``[..., 8C]``: nibble p of word c lands at column ``8c + order[p]``."""
    shifts = torch.arange(8, device=packed.device, dtype=torch.int32) * 4
    nib = (packed.unsqueeze(-1) >> shifts) & 0xF                       # [..., C, 8] by nibble position
    out = torch.empty_like(nib)
    out[..., list(order)] = nib                                        # position p -> channel order[p]
    return out.flatten(-2)


def _pack_k_natural(codes_nk: torch.Tensor) -> torch.Tensor:
    """codes ``[N, K]`` (0..15) -> reference ``qweight int32 [N, K/8]`` (nibble j of word i = K index 8i + j)."""
    n, k = codes_nk.shape
    if k % 8 != 0:
        raise ValueError(f"in_features {k} must be a multiple of 8")
    shifts = torch.arange(8, device=codes_nk.device, dtype=torch.int64) * 4
    words = (codes_nk.reshape(n, k // 8, 8).to(torch.int64) << shifts).sum(dim=-1)
    words = torch.where(words >= 2**31, words - 2**32, words)           # wrap into the int32 range
    return words.to(torch.int32)


def awq_to_reference(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor):
    """AutoAWQ GEMM tensors of one Linear -> ``(qweight [N, K/8] int32, scales [N, K/g] fp32, zeros [N, K/g] fp32)``."""
    if qweight.dtype != torch.int32 or qzeros.dtype != torch.int32:
        raise ValueError("AWQ qweight / qzeros must be int32")
    k, n8 = qweight.shape
    n = n8 * 8
    groups = qzeros.shape[0]
    if scales.shape != (groups, n) or qzeros.shape[1] != n8 or k % groups != 0:
        raise ValueError(f"inconsistent AWQ shapes: qweight {tuple(qweight.shape)}, qzeros {tuple(qzeros.shape)}, "
                         f"scales {tuple(scales.shape)}")
    codes_kn = _unpack_last_dim(qweight, AWQ_ORDER)                      # [K, N]
    zeros_gn = _unpack_last_dim(qzeros, AWQ_ORDER)                       # [K/g, N]
    return (_pack_k_natural(codes_kn.t().contiguous()), scales.t().float().contiguous(),
            zeros_gn.t().float().contiguous())


def gptq_to_reference(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                      g_idx: torch.Tensor | None = None, zero_offset: int = 1):
    """AutoGPTQ tensors of one Linear (4-bit, no activation reordering) -> the reference layout. ``zero_offset`` is
    what the format subtracts from the zero point before storing it (1 for AutoGPTQ's v1 format, 0 for v2)."""
    if qweight.dtype != torch.int32 or qzeros.dtype != torch.int32:
        raise ValueError("GPTQ qweight / qzeros must be int32")
    k8, n = qweight.shape
    k = k8 * 8
    groups = qzeros.shape[0]
    if scales.shape != (groups, n) or qzeros.shape[1] * 8 != n or k % groups != 0:
        raise ValueError(f"inconsistent GPTQ shapes: qweight {tuple(qweight.shape)}, qzeros {tuple(qzeros.shape)}, "
                         f"scales {tuple(scales.shape)}")
    if g_idx is not None:
        expect = torch.arange(k, device=g_idx.device) // (k // groups)
        if not torch.equal(g_idx.to(expect.dtype), expect):
            raise ValueError("GPTQ act-order (non-trivial g_idx) checkpoints are not supported")
    codes_nk = _unpack_last_dim(qweight.t().contiguous())                # [N, K]: natural order along K
    zeros_gn = _unpack_last_dim(qzeros) + zero_offset                    # [K/g, N]
    return (_pack_k_natural(codes_nk), scales.t().float().contiguous(), zeros_gn.t().float().contiguous())

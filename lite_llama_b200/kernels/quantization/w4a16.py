"""W4A16 GEMM (reference: lite_llama/kernels/quantization/w4a16.py:152-207)."""

from __future__ import annotations

import torch

from .._common import check, counters, lib, ptr, require_cuda, stream_handle, workspace

_Q_INT4 = 0


@torch.no_grad()
def w4a16_matmul(
    x: torch.Tensor,
    qweight: torch.Tensor,
    scales: torch.Tensor,
    zeros: torch.Tensor,
    *,
    group_size: int = 128,
    bias: torch.Tensor | None = None,
) -> torch.Tensor:
    """``x @ dequant(qweight).T (+ bias)``; ``qweight [N, K/8]`` int32 (nibble j of word i = K index
    8i+j), ``scales``/``zeros`` fp32 ``[N, K/group_size]``. Raises like the reference on bad dtypes."""
    if x.dtype != torch.float16:
        raise ValueError(f"w4a16 activations must be fp16, got {x.dtype}")
    if qweight.dtype != torch.int32:
        raise ValueError(f"qweight must be int32 (packed int4), got {qweight.dtype}")
    require_cuda(x, qweight, scales, zeros, bias)
    n, k_packed = qweight.shape
    k = k_packed * 8
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    if k % group_size != 0:
        raise ValueError(f"K ({k}) must be a multiple of group_size ({group_size})")
    if scales.dtype != torch.float32 or zeros.dtype != torch.float32:
        scales, zeros = scales.float(), zeros.float()
    if zeros.stride() != scales.stride():
        zeros = zeros.contiguous()
        scales = scales.contiguous()
    if qweight.stride(-1) != 1:
        qweight = qweight.contiguous()
    leading = x.shape[:-1]
    a = x.reshape(-1, k)
    if a.stride(-1) != 1 or a.stride(0) % 8 != 0 or a.data_ptr() % 16 != 0:
        a = a.contiguous()
    m = a.shape[0]
    out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    bias_f32 = 0
    if bias is not None:
        if bias.dtype == torch.float32:
            bias_f32 = 1
        elif bias.dtype != torch.float16:
            bias = bias.float()
            bias_f32 = 1
        bias = bias.contiguous()
    L = lib()
    nbytes = L.b2l_quant_gemm_workspace_bytes(_Q_INT4, m, n, k, int(group_size))
    ws = workspace(nbytes, x.device)
    cnt = counters(x.device)
    check(
        L.b2l_w4a16_matmul(
            out.data_ptr(), a.data_ptr(), qweight.data_ptr(), scales.data_ptr(), zeros.data_ptr(), ptr(bias),
            bias_f32, m, n, k, int(group_size), a.stride(0), qweight.stride(0), scales.stride(0), scales.stride(1),
            out.stride(0), ptr(ws), nbytes, cnt.data_ptr(), cnt.numel(), stream_handle(),
        )
    )
    return out.reshape(*leading, n)

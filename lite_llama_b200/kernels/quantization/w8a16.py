"""W8A16 GEMM (reference: lite_llama/kernels/quantization/w8a16.py:155-216)."""

from __future__ import annotations

import torch

from .._common import check, counters, lib, ptr, require_cuda, stream_handle, zeroed_workspace

_Q_FP8, _Q_INT8 = 1, 2


@torch.no_grad()
def w8a16_matmul(
    x: torch.Tensor,
    qweight: torch.Tensor,
    scales: torch.Tensor,
    *,
    group_n: int,
    group_k: int,
    bias: torch.Tensor | None = None,
) -> torch.Tensor:
    """``x @ dequant(qweight).T (+ bias)``; ``qweight [N, K]`` uint8 (fp8-e4m3 bits) or int8,
    ``scales [ceil(N/group_n), ceil(K/group_k)]`` fp32."""
    is_fp8 = qweight.dtype == torch.uint8
    if not is_fp8 and qweight.dtype != torch.int8:
        raise ValueError(f"qweight must be uint8 (fp8) or int8, got {qweight.dtype}")
    if x.dtype != torch.float16:
        raise ValueError(f"w8a16 activations must be fp16, got {x.dtype}")
    if qweight.stride(-1) != 1:
        raise ValueError("qweight last dimension must be contiguous")
    require_cuda(x, qweight, scales, bias)
    n, k = qweight.shape
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    if group_k % 128 != 0 and group_k < k:
        raise ValueError(f"group_k ({group_k}) must be a multiple of 128 unless it covers K")
    if scales.dtype != torch.float32:
        scales = scales.float()
    if scales.dim() == 1:
        scales = scales.unsqueeze(-1)
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
    kind = _Q_FP8 if is_fp8 else _Q_INT8
    nbytes = L.b2l_quant_gemm_workspace_bytes(kind, m, n, k, int(group_k))
    ws = zeroed_workspace(nbytes, x.device)
    cnt = counters(x.device)
    check(
        L.b2l_w8a16_matmul(
            out.data_ptr(), a.data_ptr(), qweight.data_ptr(), 1 if is_fp8 else 0, scales.data_ptr(), ptr(bias),
            bias_f32, m, n, k, int(group_n), int(group_k), a.stride(0), qweight.stride(0), scales.stride(0),
            scales.stride(1), out.stride(0), ptr(ws), nbytes, cnt.data_ptr(), cnt.numel(), stream_handle(),
        )
    )
    return out.reshape(*leading, n)

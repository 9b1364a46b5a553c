"""SmoothQuant W8A8 GEMM (reference: lite_llama/kernels/quantization/w8a8.py:151-217)."""

from __future__ import annotations

import torch

from .._common import check, counters, dtype_code, lib, ptr, require_cuda, stream_handle, workspace


@torch.no_grad()
def quantize_activations_int8(a: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-token dynamic int8 quantisation: ``scale = absmax/127``, ``q = trunc(x / scale)``
    (reference kernel: w8a8.py:35-68). Returns ``(q [M, K] int8, scales [M] fp32)``."""
    require_cuda(a)
    m, k = a.shape
    q = torch.empty((m, k), dtype=torch.int8, device=a.device)
    s = torch.empty((m,), dtype=torch.float32, device=a.device)
    check(
        lib().b2l_quantize_activations_int8(
            q.data_ptr(), s.data_ptr(), a.data_ptr(), m, k, a.stride(0), q.stride(0), dtype_code(a, "x"),
            stream_handle(),
        )
    )
    return q, s


@torch.no_grad()
def smoothquant_matmul_prequant(
    qa: torch.Tensor,
    a_scales: torch.Tensor,
    qweight: torch.Tensor,
    weight_scales: torch.Tensor,
    *,
    bias: torch.Tensor | None = None,
    out_dtype: torch.dtype = torch.float16,
) -> torch.Tensor:
    """The int8 GEMM half of :func:`smoothquant_matmul` (w8a8.py:75-130) on activations that are already
    quantised (``qa [M, K] int8``, ``a_scales [M] fp32``): ``(qa @ qweight.T) * a_scales[:, None] * weight_scales
    (+ bias)`` in ``out_dtype``."""
    require_cuda(qa, a_scales, qweight, weight_scales, bias)
    if qa.dtype != torch.int8 or qweight.dtype != torch.int8:
        raise ValueError("smoothquant_matmul_prequant: activations and weights must be int8")
    if out_dtype not in (torch.float16, torch.bfloat16):
        raise ValueError(f"smoothquant_matmul_prequant: unsupported output dtype {out_dtype}")
    n, k = qweight.shape
    m = qa.shape[0]
    if qa.dim() != 2 or qa.shape[1] != k or qa.stride(1) != 1:
        raise ValueError(f"qa must be [M, {k}] with a contiguous last dimension")
    if k % 16 != 0:
        raise ValueError(f"K ({k}) must be a multiple of 16 for the int8 tensor-core path")
    if weight_scales.dim() > 1:
        weight_scales = weight_scales.squeeze(-1)
    weight_scales = weight_scales.float().contiguous()
    a_scales = a_scales.float().contiguous()
    if qweight.stride(-1) != 1 or qweight.stride(0) % 16 != 0:
        qweight = qweight.contiguous()
    if bias is not None:
        bias = bias.float().contiguous()
    out = torch.empty((m, n), dtype=out_dtype, device=qa.device)
    L = lib()
    nbytes = L.b2l_linear_workspace_bytes(m, n, k, 1)
    ws = workspace(nbytes, qa.device)
    cnt = counters(qa.device)
    check(
        L.b2l_w8a8_gemm(
            out.data_ptr(), qa.data_ptr(), qweight.data_ptr(), a_scales.data_ptr(), weight_scales.data_ptr(),
            ptr(bias), m, n, k, qa.stride(0), qweight.stride(0), out.stride(0), ptr(ws), nbytes, cnt.data_ptr(),
            cnt.numel(), dtype_code(out, "out"), stream_handle(),
        )
    )
    return out


@torch.no_grad()
def smoothquant_matmul(
    x: torch.Tensor,
    qweight: torch.Tensor,
    weight_scales: torch.Tensor,
    *,
    bias: torch.Tensor | None = None,
) -> torch.Tensor:
    """``x @ dequant(qweight).T (+ bias)`` with dynamic per-token int8 activations."""
    if x.dtype != torch.float16:
        raise ValueError(f"smoothquant activations must be fp16, got {x.dtype}")
    if qweight.dtype != torch.int8:
        raise ValueError(f"qweight must be int8, got {qweight.dtype}")
    require_cuda(x, qweight, weight_scales, bias)
    n, k = qweight.shape
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    if k % 16 != 0:
        raise ValueError(f"K ({k}) must be a multiple of 16 for the int8 tensor-core path")
    leading = x.shape[:-1]
    a = x.reshape(-1, k)
    if a.stride(-1) != 1:
        a = a.contiguous()
    qa, a_scales = quantize_activations_int8(a)
    out = smoothquant_matmul_prequant(qa, a_scales, qweight, weight_scales, bias=bias, out_dtype=x.dtype)
    return out.reshape(*leading, n)

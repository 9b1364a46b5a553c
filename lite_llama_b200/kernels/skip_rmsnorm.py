"""Fused residual-add + RMSNorm (reference: lite_llama/kernels/skip_rmsnorm.py:192-234)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, lib, ptr, require_cuda, stream_handle


@torch.no_grad()
def skip_rmsnorm(X, residual, weight, eps=1e-5):
    """``(Y, residual)``: ``residual += X`` in place, ``Y = rmsnorm(residual) * weight``.

    With ``residual=None`` this is plain RMSNorm and ``X`` is echoed back as the second value,
    exactly like the reference (skip_rmsnorm.py:219,234). The last dim is normalised; any leading
    shape is accepted (the per-head q/k norm passes ``[tokens, heads, head_dim]``).
    """
    require_cuda(X, residual, weight)
    orig_shape = X.shape
    n = orig_shape[-1]
    x2 = X.contiguous().view(-1, n)
    code = dtype_code(x2, "X")
    if weight.dtype != x2.dtype:
        raise ValueError(f"weight dtype {weight.dtype} must match X dtype {x2.dtype}")
    w = weight.contiguous()
    y = torch.empty_like(x2)
    r2 = None
    if residual is not None:
        if residual.dtype != x2.dtype:
            raise ValueError(f"residual dtype {residual.dtype} must match X dtype {x2.dtype}")
        r2 = residual.contiguous().view(-1, n)  # aliases the caller's residual when it is contiguous
    check(
        lib().b2l_skip_rmsnorm(
            y.data_ptr(), x2.data_ptr(), ptr(r2), w.data_ptr(), x2.shape[0], n, x2.stride(0),
            r2.stride(0) if r2 is not None else 0, y.stride(0), float(eps), code, stream_handle(),
        )
    )
    if r2 is not None:
        return y.view(orig_shape), r2.view(orig_shape)
    return y.view(orig_shape), x2.view(orig_shape)


def can_fuse_int8_quant(n: int) -> bool:
    """Shapes the fused norm + int8 quantisation kernel takes (16-byte rows kept in registers)."""
    return n % 8 == 0 and n <= 16384


@torch.no_grad()
def skip_rmsnorm_quant_int8(X, residual, weight, eps=1e-5):
    """``skip_rmsnorm`` whose output goes straight into a SmoothQuant GEMM: returns ``(q int8 [rows, n],
    scales fp32 [rows], residual)`` -- the per-token dynamic quantisation of the (fp16-rounded) norm output
    (reference: skip_rmsnorm.py:192-234 then quantization/w8a8.py:35-68), bit-identical to calling the two ops,
    in one launch and without writing the fp16 activations."""
    require_cuda(X, residual, weight)
    orig_shape = X.shape
    n = orig_shape[-1]
    x2 = X.contiguous().view(-1, n)
    code = dtype_code(x2, "X")
    if weight.dtype != x2.dtype:
        raise ValueError(f"weight dtype {weight.dtype} must match X dtype {x2.dtype}")
    if not can_fuse_int8_quant(n):
        raise ValueError(f"skip_rmsnorm_quant_int8 needs n % 8 == 0 and n <= 16384, got {n}")
    w = weight.contiguous()
    r2 = None
    if residual is not None:
        if residual.dtype != x2.dtype:
            raise ValueError(f"residual dtype {residual.dtype} must match X dtype {x2.dtype}")
        r2 = residual.contiguous().view(-1, n)
    q = torch.empty((x2.shape[0], n), dtype=torch.int8, device=x2.device)
    s = torch.empty((x2.shape[0],), dtype=torch.float32, device=x2.device)
    check(
        lib().b2l_skip_rmsnorm_quant_int8(
            q.data_ptr(), s.data_ptr(), x2.data_ptr(), ptr(r2), w.data_ptr(), x2.shape[0], n, x2.stride(0),
            r2.stride(0) if r2 is not None else 0, q.stride(0), float(eps), code, stream_handle(),
        )
    )
    return q, s, (r2.view(orig_shape) if r2 is not None else x2.view(orig_shape))

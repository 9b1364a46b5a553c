"""Dense projection ``x @ w.T (+ bias)`` on tcgen05.

Stands in for the reference's ``F.linear`` call sites
(lite_llama/models/quantization/methods/unquantized.py:21-22, lite_llama/models/base.py:489).
"""

from __future__ import annotations

import torch

from ._common import check, counters, dtype_code, lib, ptr, require_cuda, stream_handle, workspace


@torch.no_grad()
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """``[..., K] @ [N, K].T -> [..., N]`` in ``x``'s dtype, fp32 accumulation."""
    require_cuda(x, weight, bias)
    code = dtype_code(x, "x")
    if weight.dtype != x.dtype:
        raise ValueError(f"weight dtype {weight.dtype} must match activation dtype {x.dtype}")
    n, k = weight.shape
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    if weight.stride(-1) != 1:
        weight = weight.contiguous()
    leading = x.shape[:-1]
    a = x.reshape(-1, k)
    if a.stride(-1) != 1 or a.stride(0) % 8 != 0 or a.data_ptr() % 16 != 0:
        a = a.contiguous()
    m = a.shape[0]
    out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    L = lib()
    nbytes = L.b2l_linear_workspace_bytes(m, n, k, 2)
    ws = workspace(nbytes, x.device)
    cnt = counters(x.device)
    check(
        L.b2l_linear(
            out.data_ptr(), a.data_ptr(), weight.data_ptr(), ptr(bias), m, n, k, a.stride(0), weight.stride(0),
            out.stride(0), ptr(ws), nbytes, cnt.data_ptr(), cnt.numel(), code, stream_handle(),
        )
    )
    return out.reshape(*leading, n)

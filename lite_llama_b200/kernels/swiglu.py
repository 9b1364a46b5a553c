"""SiLU(gate) * up (reference: lite_llama/kernels/swiglu.py:45-65)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, lib, require_cuda, stream_handle


@torch.no_grad()
def swiglu_forward(a, b):
    require_cuda(a, b)
    ori_shape = a.shape
    n_cols = ori_shape[-1]
    a2 = a.reshape(-1, n_cols)
    b2 = b.reshape(-1, n_cols)
    if a2.stride(-1) != 1:
        a2 = a2.contiguous()
    if b2.stride(-1) != 1:
        b2 = b2.contiguous()
    code = dtype_code(a2, "a")
    c = torch.empty((a2.shape[0], n_cols), dtype=a.dtype, device=a.device)
    check(
        lib().b2l_swiglu(
            c.data_ptr(), a2.data_ptr(), b2.data_ptr(), a2.shape[0], n_cols, a2.stride(0), b2.stride(0), c.stride(0),
            code, stream_handle(),
        )
    )
    return c.view(*ori_shape)

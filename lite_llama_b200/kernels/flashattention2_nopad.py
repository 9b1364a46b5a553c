"""Causal varlen prefill attention (reference: lite_llama/kernels/flashattention2_nopad.py:175-231)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, index_width, lib, require_cuda, stream_handle


@torch.no_grad()
def flash_attention2_no_pad(q, k, v, sm_scale, b_start_loc, b_seq_len, max_seq_len):
    """``q [T, Hq, D]``, ``k``/``v`` ``[T, Hkv, D]``; ``sm_scale`` already multiplied by log2(e).

    The reference wraps its launcher in ``custom_fwd(cast_inputs=torch.float16)`` (:176); autocast
    is never active on the inference path, so inputs are used in the dtype they arrive in.
    """
    require_cuda(q, k, v, b_start_loc, b_seq_len)
    code = dtype_code(q, "q")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("flash_attention2_no_pad: q, k and v must share one dtype")
    if q.stride(2) != 1:
        q = q.contiguous()
    if k.stride(2) != 1:
        k = k.contiguous()
    if v.stride(2) != 1:
        v = v.contiguous()
    output = torch.empty_like(q)
    batchs = b_seq_len.shape[0]
    n_heads, head_dim = q.shape[1], q.shape[2]
    check(
        lib().b2l_flash_attention2_no_pad(
            output.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(),
            b_start_loc.data_ptr(), index_width(b_start_loc, "b_start_loc"),
            b_seq_len.data_ptr(), index_width(b_seq_len, "b_seq_len"),
            batchs, n_heads, k.shape[1], head_dim, int(max_seq_len), int(q.shape[0]), float(sm_scale),
            q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1),
            output.stride(0), output.stride(1), code, stream_handle(),
        )
    )
    return output

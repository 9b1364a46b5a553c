"""Half-split RoPE on q and k, in place (reference: lite_llama/kernels/rope_emb.py:86-134)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, index_width, lib, require_cuda, stream_handle


@torch.no_grad()
def rope_emb_forward(q, k, cos, sin, batch_size, seq_len):
    """Rotate ``q [B*S, Hq, D]`` and ``k [B*S, Hkv, D]`` by ``cos/sin [B, S, D]`` (first D/2 read).

    Like the reference, non-contiguous inputs are made contiguous first (so a strided ``k`` view is
    copied, rope_emb.py:106-107) and the returned tensors are the ones to use.
    """
    require_cuda(q, k, cos, sin)
    n, n_qh, hd = q.shape
    _, n_kh, _ = k.shape
    assert batch_size * seq_len == n
    code = dtype_code(q, "q")
    q = q.contiguous()
    k = k.contiguous()
    cos = cos.contiguous()
    sin = sin.contiguous()
    if cos.dtype != q.dtype:  # the reference computes in the cos/sin dtype; the caller casts to x.dtype
        cos = cos.to(q.dtype)
        sin = sin.to(q.dtype)
    check(
        lib().b2l_rope_emb(
            q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), batch_size, seq_len, n_qh, n_kh, hd,
            q.stride(0), q.stride(1), k.stride(0), k.stride(1), cos.stride(0), cos.stride(1), code, stream_handle(),
        )
    )
    return q, k


@torch.no_grad()
def rope_emb_inplace(q, k, cos, sin, batch_size, seq_len):
    """Stride-aware variant used by the host layer: rotates ``q``/``k`` where they lie (e.g. the q and
    k column blocks of one fused QKV projection output), so no ``.contiguous()`` copy of ``k`` is made
    (the copy the reference pays at rope_emb.py:107). Only the head_dim axis must be contiguous."""
    require_cuda(q, k, cos, sin)
    n, n_qh, hd = q.shape
    _, n_kh, _ = k.shape
    assert batch_size * seq_len == n
    if q.stride(2) != 1 or k.stride(2) != 1 or cos.stride(-1) != 1 or sin.stride(-1) != 1:
        raise ValueError("rope_emb_inplace: innermost dimension must be contiguous")
    if cos.dtype != q.dtype or sin.dtype != q.dtype:
        raise ValueError("rope_emb_inplace: cos/sin must already be in the activation dtype")
    check(
        lib().b2l_rope_emb(
            q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), batch_size, seq_len, n_qh, n_kh, hd,
            q.stride(0), q.stride(1), k.stride(0), k.stride(1), cos.stride(0), cos.stride(1),
            dtype_code(q, "q"), stream_handle(),
        )
    )
    return q, k


@torch.no_grad()
def rope_kv_write(q, k, v, cos, sin, select_index, kv_buffer, batch_size, seq_len):
    """RoPE on ``q``/``k`` (in place, where they lie) + the K|V cache write in one launch: what the reference does
    with ``rope_emb_forward`` -> ``torch.cat([xk, xv])`` -> ``update_kv_buffer`` (models/base.py:76-79,228).
    ``kv_buffer [max_tokens, 2*Hkv, D]`` receives the rotated ``k`` rows and the ``v`` rows at ``select_index``."""
    require_cuda(q, k, v, cos, sin, select_index, kv_buffer)
    n, n_qh, hd = q.shape
    n_kh = k.shape[1]
    assert batch_size * seq_len == n and v.shape[1] == n_kh and kv_buffer.shape[1] == 2 * n_kh
    if any(t.stride(-1) != 1 for t in (q, k, v, cos, sin, kv_buffer)):
        raise ValueError("rope_kv_write: innermost dimension must be contiguous")
    if cos.dtype != q.dtype or sin.dtype != q.dtype or kv_buffer.dtype != q.dtype:
        raise ValueError("rope_kv_write: cos/sin/kv_buffer must be in the activation dtype")
    check(
        lib().b2l_rope_kv_write(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), cos.data_ptr(), sin.data_ptr(), select_index.data_ptr(),
            index_width(select_index, "select_index"), kv_buffer.data_ptr(), batch_size, seq_len, n_qh, n_kh, hd,
            q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), cos.stride(0), cos.stride(1),
            kv_buffer.stride(0), kv_buffer.stride(1), dtype_code(q, "q"), stream_handle(),
        )
    )
    return q, k

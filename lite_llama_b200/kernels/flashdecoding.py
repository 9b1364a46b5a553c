"""Decode attention over the token-granular paged KV cache
(reference: lite_llama/kernels/flashdecoding.py:316-380)."""

from __future__ import annotations

import os

import torch

from ._common import check, dtype_code, index_width, lib, ptr, require_cuda, stream_handle, workspace, counters

# Folding the split-KV merge into the last-arriving split CTA (tickets) removes a launch but measured SLOWER at decode
# sizes (24.8 us vs 11.3 + ~3 us: one CTA merges 4096 outputs serially at L2 latency, profiles/r2f_*): off by default.
_FOLD = os.environ.get("B2L_FD_FOLD", "0") != "0"


@torch.no_grad()
def flash_decoding(
    q, k_cache, v_cache, qk_scale, b_req_tokens_table, b_req_idx, b_seq_len, max_actual_seq_len
):
    """One new query per sequence against its cached history.

    Args mirror the reference: ``q [B, Hq, D]``; ``k_cache``/``v_cache`` ``[max_tokens, Hkv, D]``
    (usually strided views of one ``[max_tokens, 2*Hkv, D]`` pool); ``qk_scale = 1/sqrt(D)``;
    ``b_req_tokens_table [max_requests, max_seq_len]`` int32; ``b_req_idx``/``b_seq_len`` ``[B]``
    int32 or int64; ``max_actual_seq_len`` a Python int (sizes the split grid; graph-safe).
    """
    require_cuda(q, k_cache, v_cache, b_req_tokens_table, b_req_idx, b_seq_len)
    assert q.shape[-1] == k_cache.shape[-1] == v_cache.shape[-1]
    batch, num_heads, head_dim = q.shape
    num_kv_heads = k_cache.shape[1]
    code = dtype_code(q, "q")
    if k_cache.dtype != q.dtype or v_cache.dtype != q.dtype:
        raise ValueError("flash_decoding: q, k_cache and v_cache must share one dtype")
    if b_req_tokens_table.dtype != torch.int32:
        raise ValueError(f"b_req_tokens_table must be int32, got {b_req_tokens_table.dtype}")
    if q.stride(2) != 1:
        q = q.contiguous()
    if k_cache.stride(2) != 1 or v_cache.stride(2) != 1:
        raise ValueError("flash_decoding: K/V cache head_dim must be contiguous")
    out = torch.empty_like(q)
    max_len = int(max_actual_seq_len)
    nbytes = lib().b2l_flash_decoding_workspace_bytes(batch, num_heads, num_kv_heads, head_dim, max_len)
    ws = workspace(nbytes, q.device)
    # B2L_FD_FOLD=1: zeroed, self-resetting tickets; the split CTA that finishes a (row, head block) last merges the
    # splits itself (no second launch)
    cnt = counters(q.device) if _FOLD else None
    check(
        lib().b2l_flash_decoding(
            out.data_ptr(), q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), b_req_tokens_table.data_ptr(),
            b_req_idx.data_ptr(), index_width(b_req_idx, "b_req_idx"),
            b_seq_len.data_ptr(), index_width(b_seq_len, "b_seq_len"),
            batch, num_heads, num_kv_heads, head_dim, max_len, float(qk_scale),
            q.stride(0), q.stride(1), k_cache.stride(0), k_cache.stride(1), v_cache.stride(0), v_cache.stride(1),
            b_req_tokens_table.stride(0), b_req_tokens_table.stride(1), out.stride(0), out.stride(1),
            ptr(ws), nbytes, ptr(cnt), cnt.numel() if cnt is not None else 0, code, stream_handle(),
        )
    )
    return out

"""Scatter new K/V rows into the paged pool (reference: lite_llama/kernels/update_kv_buffer.py:55-89)."""

from __future__ import annotations

import torch

from ._common import check, index_width, lib, require_cuda, stream_handle


@torch.no_grad()
def update_kv_buffer(KV_Values, Select_Index, KV_Buffer):
    """``KV_Buffer[Select_Index[i]] = KV_Values[i]`` for ``[tokens, 2*Hkv, D]`` rows."""
    require_cuda(KV_Values, Select_Index, KV_Buffer)
    assert KV_Values.shape[1] == KV_Buffer.shape[1] and KV_Values.shape[2] == KV_Buffer.shape[2]
    if KV_Values.element_size() != 2 or KV_Buffer.element_size() != 2:
        raise ValueError("update_kv_buffer expects 2-byte (fp16/bf16) tensors")
    if KV_Values.stride(2) != 1 or KV_Buffer.stride(2) != 1:
        raise ValueError("update_kv_buffer: innermost dimension must be contiguous")
    tokens = Select_Index.shape[0]
    check(
        lib().b2l_update_kv_buffer(
            KV_Values.data_ptr(), None, Select_Index.data_ptr(), index_width(Select_Index, "Select_Index"),
            KV_Buffer.data_ptr(), tokens, KV_Values.shape[1], 0, KV_Values.shape[2],
            KV_Values.stride(0), KV_Values.stride(1), 0, 0, KV_Buffer.stride(0), KV_Buffer.stride(1),
            stream_handle(),
        )
    )


@torch.no_grad()
def update_kv_buffer_split(K_Values, V_Values, Select_Index, KV_Buffer):
    """Same scatter from separate K and V row tensors (no ``torch.cat``; lite_llama/models/base.py:76)."""
    require_cuda(K_Values, V_Values, Select_Index, KV_Buffer)
    hk, hv, d = K_Values.shape[1], V_Values.shape[1], K_Values.shape[2]
    assert hk + hv == KV_Buffer.shape[1] and d == KV_Buffer.shape[2] == V_Values.shape[2]
    if K_Values.stride(2) != 1 or V_Values.stride(2) != 1 or KV_Buffer.stride(2) != 1:
        raise ValueError("update_kv_buffer: innermost dimension must be contiguous")
    check(
        lib().b2l_update_kv_buffer(
            K_Values.data_ptr(), V_Values.data_ptr(), Select_Index.data_ptr(),
            index_width(Select_Index, "Select_Index"), KV_Buffer.data_ptr(), Select_Index.shape[0], hk, hv, d,
            K_Values.stride(0), K_Values.stride(1), V_Values.stride(0), V_Values.stride(1),
            KV_Buffer.stride(0), KV_Buffer.stride(1), stream_handle(),
        )
    )

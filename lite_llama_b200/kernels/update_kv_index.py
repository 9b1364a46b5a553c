"""``table[b_req_idx[i], b_seq_len[i]-1] = select_index[i]`` (reference: kernels/update_kv_index.py:51-88)."""

from __future__ import annotations

import torch

from ._common import check, index_width, lib, require_cuda, stream_handle


@torch.no_grad()
def update_kv_index(req_to_token_indexs, b_req_idx, b_seq_len, select_index):
    require_cuda(req_to_token_indexs, b_req_idx, b_seq_len, select_index)
    assert (
        b_seq_len.shape[0] == select_index.shape[0] and b_req_idx.shape[0] == b_seq_len.shape[0]
    ), "all index vectors must have the same length"
    if req_to_token_indexs.dtype != torch.int32:
        raise ValueError(f"req_to_token_indexs must be int32, got {req_to_token_indexs.dtype}")
    check(
        lib().b2l_update_kv_index(
            req_to_token_indexs.data_ptr(), b_req_idx.data_ptr(), index_width(b_req_idx, "b_req_idx"),
            b_seq_len.data_ptr(), index_width(b_seq_len, "b_seq_len"), select_index.data_ptr(),
            index_width(select_index, "select_index"), b_seq_len.shape[0],
            req_to_token_indexs.stride(0), req_to_token_indexs.stride(1), stream_handle(),
        )
    )

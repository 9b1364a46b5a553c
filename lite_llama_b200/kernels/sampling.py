"""Greedy sampling helper: argmax over the vocabulary (reference: the temperature-0 branch of
lite_llama/engine/sampler.py, ``torch.argmax`` on ``[batch, vocab]`` logits, engine/llm_engine.py:172)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, lib, require_cuda, stream_handle


@torch.no_grad()
def argmax_rows(logits: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``logits [rows, n]`` (last dim contiguous) -> int64 ``[rows]``; ties resolve to the lowest index."""
    require_cuda(logits, out)
    if logits.dim() != 2 or logits.stride(1) != 1:
        raise ValueError("argmax_rows expects [rows, n] with a contiguous last dimension")
    rows, n = logits.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    elif out.dtype != torch.int64 or out.numel() != rows or not out.is_contiguous():
        raise ValueError("argmax_rows: out must be a contiguous int64 tensor with one entry per row")
    check(lib().b2l_argmax_rows(out.data_ptr(), logits.data_ptr(), rows, n, logits.stride(0),
                                dtype_code(logits, "logits"), stream_handle()))
    return out

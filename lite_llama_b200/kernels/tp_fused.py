"""Row-parallel epilogue: all-reduce of the partial projection output fused with the skip_rmsnorm that follows it.

Reference sequence (lite_llama/models/linear.py:160-161 -> lite_llama/models/base.py:307,312,486):
``y = all_reduce_tp(partial)`` then ``skip_rmsnorm(y, residual, weight, eps)``. Here both are one kernel over
NVLink peer memory (csrc/tp_comm.cu) whenever the communicator exists and the step is at most
``NvlinkComm.max_rows`` tokens wide; otherwise NCCL carries the all-reduce and skip_rmsnorm runs as usual.
Same return convention as :func:`skip_rmsnorm`: ``(Y, residual)`` with ``residual`` updated in place.
"""

from __future__ import annotations

import torch

from ..distributed import all_reduce_tp, get_tp_comm, get_tp_world_size
from ._common import check, dtype_code, lib, require_cuda, stream_handle
from .skip_rmsnorm import skip_rmsnorm


@torch.no_grad()
def allreduce_skip_rmsnorm(partial: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5):
    if get_tp_world_size() <= 1:
        return skip_rmsnorm(partial, residual, weight, eps)
    comm = get_tp_comm()
    n = partial.shape[-1]
    rows = partial.numel() // n
    if comm is None or not comm.fits(rows, n) or not partial.is_cuda:
        return skip_rmsnorm(all_reduce_tp(partial), residual, weight, eps)
    require_cuda(partial, residual, weight)
    if residual is None:
        raise ValueError("allreduce_skip_rmsnorm needs the residual stream (every row-parallel output joins one)")
    x2 = partial.contiguous().view(-1, n)
    r2 = residual.contiguous().view(-1, n)
    if r2.data_ptr() != residual.data_ptr():
        raise ValueError("residual must be contiguous (it is updated in place)")
    code = dtype_code(x2, "partial")
    if weight.dtype != x2.dtype or r2.dtype != x2.dtype:
        raise ValueError("partial, residual and weight must share one dtype")
    y = torch.empty_like(x2)
    check(
        lib().b2l_tp_allreduce_skip_rmsnorm(
            comm.handle, y.data_ptr(), x2.data_ptr(), r2.data_ptr(), weight.contiguous().data_ptr(), rows, n,
            x2.stride(0), r2.stride(0), y.stride(0), float(eps), code, stream_handle(),
        )
    )
    return y.view(partial.shape), r2.view(residual.shape)


@torch.no_grad()
def allreduce_nvlink(partial: torch.Tensor) -> torch.Tensor:
    """Plain SUM all-reduce through the same kernel (used by tests and the bench's latency probe)."""
    comm = get_tp_comm()
    n = partial.shape[-1]
    rows = partial.numel() // n
    if comm is None or not comm.fits(rows, n):
        return all_reduce_tp(partial)
    x2 = partial.contiguous().view(-1, n)
    y = torch.empty_like(x2)
    check(
        lib().b2l_tp_allreduce_skip_rmsnorm(
            comm.handle, y.data_ptr(), x2.data_ptr(), None, None, rows, n, x2.stride(0), 0, y.stride(0), 0.0,
            dtype_code(x2, "partial"), stream_handle(),
        )
    )
    return y.view(partial.shape)

"""Tensor-parallel process-group state (reference: lite_llama/distributed/parallel_state.py:44-230).

One process per GPU; the TP group is the world group. The hot-path collective is the in-place SUM
all-reduce of ``[tokens, hidden]`` after every row-parallel projection (models/linear.py:160-161).
``backend="gloo"`` is used by the CPU tests, ``"nccl"`` on the GPUs.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist

_TP_GROUP = None
_TP_SIZE = 1
_TP_RANK = 0
_TP_COMM = None      # NvlinkComm (peer-memory exchange for the fused all-reduce + norm kernel) or None


def divide(numerator: int, denominator: int, what: str = "value") -> int:
    """Exact division or a readable error (parallel_state.py:194-202)."""
    if numerator % denominator != 0:
        raise ValueError(f"{what} ({numerator}) is not divisible by tensor-parallel size {denominator}")
    return numerator // denominator


def init_tensor_parallel(tp_size: int | None = None, rank: int | None = None, backend: str | None = None,
                         nvlink: bool | None = None) -> None:
    """Join (or create) the process group; rank/world size default to the torchrun environment.

    On CUDA with more than one rank the NVLink peer-memory communicator of the fused
    all-reduce + skip_rmsnorm kernel is set up as well (``nvlink=False`` or env ``B2L_TP_NVLINK=0``
    keeps every exchange on NCCL); if peer mapping is impossible on this machine the failure is
    reported once and NCCL carries the exchange."""
    global _TP_GROUP, _TP_SIZE, _TP_RANK, _TP_COMM
    world = int(os.environ.get("WORLD_SIZE", "1")) if tp_size is None else tp_size
    r = int(os.environ.get("RANK", "0")) if rank is None else rank
    if world <= 1:
        _TP_GROUP, _TP_SIZE, _TP_RANK = None, 1, 0
        return
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=backend, rank=r, world_size=world, **kw)
    _TP_GROUP = dist.group.WORLD
    _TP_SIZE = dist.get_world_size()
    _TP_RANK = dist.get_rank()
    if nvlink is None:
        nvlink = os.environ.get("B2L_TP_NVLINK", "1") != "0"
    _TP_COMM = None
    if nvlink and _TP_SIZE > 1 and torch.cuda.is_available() and dist.get_backend() == "nccl":
        from .nvlink_comm import NvlinkComm

        ok = torch.ones(1, device="cuda")
        try:
            comm = NvlinkComm(_TP_RANK, _TP_SIZE, _TP_GROUP, torch.device("cuda", torch.cuda.current_device()))
        except Exception as ex:  # no peer access on this machine: say so, NCCL carries the exchange
            comm = None
            ok.zero_()
            print(f"[lite_llama_b200] rank {_TP_RANK}: NVLink communicator unavailable ({ex}); using NCCL", flush=True)
        # all ranks must take the same path
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=_TP_GROUP)
        _TP_COMM = comm if bool(ok.item() > 0) else None


def destroy_tensor_parallel() -> None:
    global _TP_GROUP, _TP_SIZE, _TP_RANK, _TP_COMM
    if _TP_COMM is not None:
        _TP_COMM.close()
    _TP_COMM = None
    if dist.is_initialized():
        dist.destroy_process_group()
    _TP_GROUP, _TP_SIZE, _TP_RANK = None, 1, 0


def get_tp_comm():
    """The NVLink peer-memory communicator, or ``None`` (single rank, CPU/gloo, or unavailable)."""
    return _TP_COMM


def get_tp_world_size() -> int:
    return _TP_SIZE


def get_tp_rank() -> int:
    return _TP_RANK


def all_reduce_tp(x: torch.Tensor) -> torch.Tensor:
    """In-place SUM across the TP group; identity at tp=1 (parallel_state.py:208-213)."""
    if _TP_SIZE > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=_TP_GROUP)
    return x


def all_reduce_min(value: int, device: str | torch.device = "cpu") -> int:
    """Agree on the smallest per-rank value (KV-cache size; parallel_state.py:216-230)."""
    if _TP_SIZE <= 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=_TP_GROUP)
    return int(t.item())

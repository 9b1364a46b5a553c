"""Shared plumbing for the ctypes shims: dtype codes, stream handle, scratch buffers."""

from __future__ import annotations

import torch

from .. import _lib

F16, BF16 = 0, 1

_DTYPE_CODE = {torch.float16: F16, torch.bfloat16: BF16}


def lib():
    return _lib.load()


def check(rc: int) -> None:
    _lib.check(rc)


def dtype_code(t: torch.Tensor, what: str) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise ValueError(f"{what} must be float16 or bfloat16, got {t.dtype}") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError(
                "lite_llama_b200 kernels run on CUDA tensors only (there is no CPU fallback); "
                f"got a tensor on {t.device}"
            )


def stream_handle() -> int:
    return torch.cuda.current_stream().cuda_stream


def index_width(t: torch.Tensor, what: str) -> int:
    if t.dtype == torch.int32:
        return 4
    if t.dtype == torch.int64:
        return 8
    raise ValueError(f"{what} must be int32 or int64, got {t.dtype}")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# Split-K tickets: one persistent zeroed int32 buffer per device. Kernels leave it zeroed, and all
# launches are serialised on the current stream (the reference drives one stream per process too).
_COUNTERS: dict[int, torch.Tensor] = {}
_NUM_COUNTERS = 16384


def counters(device: torch.device) -> torch.Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    buf = _COUNTERS.get(idx)
    if buf is None:
        buf = torch.zeros(_NUM_COUNTERS, dtype=torch.int32, device=device)
        _COUNTERS[idx] = buf
    return buf


# Stream-K partial sums of the quantised GEMMs travel through a workspace with flag-in-data words (csrc/gemm_quant.cu):
# the buffer must be all-zero when a launch starts and every launch leaves it all-zero, so it is persistent and zeroed
# once -- same contract (and same one-stream assumption) as the ticket counters above. A buffer that has to grow is
# replaced, never freed: a captured CUDA graph may still hold its address.
_ZEROED: dict[int, torch.Tensor] = {}
_ZEROED_RETIRED: list[torch.Tensor] = []


def zeroed_workspace(nbytes: int, device: torch.device) -> torch.Tensor | None:
    if nbytes <= 0:
        return None
    idx = device.index if device.index is not None else torch.cuda.current_device()
    buf = _ZEROED.get(idx)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ZEROED_RETIRED.append(buf)
        size = max(int(nbytes), 2 * buf.numel() if buf is not None else 0, 1 << 24)
        buf = torch.zeros(size, dtype=torch.uint8, device=device)
        _ZEROED[idx] = buf
    return buf


def workspace(nbytes: int, device: torch.device) -> torch.Tensor | None:
    """Scratch allocated through torch's caching allocator (graph-safe, like the reference's
    ``torch.empty`` scratch in flashdecoding.py:350-358)."""
    if nbytes <= 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)

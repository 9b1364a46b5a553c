"""NVLink peer-memory communicator for the fused all-reduce + skip_rmsnorm kernel (csrc/tp_comm.cu).

One process per GPU. Each rank allocates its receive buffer inside ``libb200lite.so`` (plain ``cudaMalloc``),
exports a CUDA IPC handle, the handles are all-gathered over the existing ``torch.distributed`` group
(plumbing), and every rank maps its peers' buffers. After that the hot path is a single kernel launch per
(row-parallel projection -> next norm) pair with no library collective on it.
"""

from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from .. import _lib

MAX_ROWS = 256          # decode / continuous-batching step widths; wider (prefill) batches use NCCL
MAX_ROW_ELEMS = 8192    # hidden sizes up to 8192


class NvlinkComm:
    def __init__(self, rank: int, world: int, group, device: torch.device,
                 max_rows: int = MAX_ROWS, max_row_elems: int = MAX_ROW_ELEMS) -> None:
        lib = _lib.load()
        self.rank, self.world, self.max_rows, self.max_row_elems = rank, world, max_rows, max_row_elems
        self.handle = lib.b2l_tp_comm_create(rank, world, max_rows, max_row_elems)
        if not self.handle:
            raise RuntimeError(f"b2l_tp_comm_create failed: {_lib.last_error()}")
        hb = lib.b2l_tp_comm_handle_bytes()
        buf = (ctypes.c_ubyte * hb)()
        _lib.check(lib.b2l_tp_comm_get_handle(self.handle, ctypes.cast(buf, ctypes.c_void_p), hb))
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        allh = bytes(torch.cat(gathered).cpu().tolist())
        cbuf = (ctypes.c_ubyte * len(allh)).from_buffer_copy(allh)
        _lib.check(lib.b2l_tp_comm_open(self.handle, ctypes.cast(cbuf, ctypes.c_void_p), len(allh)))
        # nobody may push before every rank has mapped every buffer
        dist.barrier(group=group)

    def fits(self, rows: int, n: int) -> bool:
        return 0 < rows <= self.max_rows and n % 2 == 0 and n <= self.max_row_elems

    def close(self) -> None:
        if self.handle:
            _lib.load().b2l_tp_comm_destroy(self.handle)
            self.handle = 0

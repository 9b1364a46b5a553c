"""Token-granular KV pool and its reference-counted row allocator.

Keeps the layout and the allocator surface the reference fixes as part of its executor API
(lite_llama/executor/kv_cache_manager.py:136-372): one tensor per layer,
``[max_tokens, 2 * num_kv_heads, head_dim]`` with the K heads first; ``kv_mem_use_state[i]`` is the
reference count of row ``i`` (free at zero) and ``can_use_mem_size`` the number of free rows.

Allocation strategies, cheapest first (kv_cache_manager.py:270-300):
  * a bump cursor while the pool is append-only -- what every ``generate()`` call sees, because it opens
    with ``free_all`` -- which costs no device read at all;
  * the first run of ``need_size`` consecutive free rows;
  * any ``need_size`` free rows.
Rows are released per reference (``release_ref`` / ``free``), all at once (``free_all``), or handed to an
external owner for good (``claim``: the slot layout of continuous batching).
"""

from __future__ import annotations

import torch


class KVCacheManager:
    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, gpu_num_blocks: int,
                 block_size: int = 1, dtype: torch.dtype = torch.float16, device: str = "cuda") -> None:
        if block_size != 1:
            raise ValueError("token attention uses block_size == 1")
        self.num_layers = num_layers
        self.num_kv_heads = num_kv_heads
        self.head_dim = head_dim
        self.gpu_num_blocks = gpu_num_blocks
        self.block_size = block_size
        self.max_num_tokens = gpu_num_blocks * block_size
        self.dtype = dtype
        self.device = device
        self.can_use_mem_size = self.max_num_tokens
        self.kv_mem_pos_indexs = torch.arange(self.max_num_tokens, dtype=torch.long, device=device)
        self.kv_mem_pos_indexs_int32 = self.kv_mem_pos_indexs.to(torch.int32)
        self.kv_mem_use_state = torch.zeros(self.max_num_tokens, dtype=torch.int32, device=device)
        self._bump_cursor = 0
        self._bump_is_exact = True
        self.gpu_kv_buffer = [
            torch.empty((self.max_num_tokens, 2 * num_kv_heads, head_dim), dtype=dtype, device=device)
            for _ in range(num_layers)
        ]

    # ------------------------------------------------------------ allocation #
    def _free_rows(self) -> torch.Tensor:
        return (self.kv_mem_use_state == 0).nonzero().view(-1)

    @torch.no_grad()
    def alloc_kvcache(self, need_size: int) -> torch.Tensor | None:
        """Any ``need_size`` free rows (lowest indices first), or ``None`` -- and then nothing changes."""
        if need_size > self.can_use_mem_size:
            return None
        rows = self._free_rows()[:need_size]
        self.add_ref(rows)
        return rows

    @torch.no_grad()
    def alloc_contiguous_kvcache(self, need_size: int):
        """``(rows, start, end)`` for the first run of ``need_size`` consecutive free rows, else ``None``."""
        if need_size > self.can_use_mem_size or need_size <= 0:
            return None
        free = self._free_rows()
        if free.numel() < need_size:
            return None
        # window j of the sorted free list spans free[j] .. free[j + need - 1]; it is a run of consecutive
        # rows exactly when the two ends are need - 1 apart
        span = free[need_size - 1:] - free[: free.numel() - need_size + 1]
        hits = (span == need_size - 1).nonzero().view(-1)
        if hits.numel() == 0:
            return None
        start = int(free[hits[0]].item())
        rows = self.kv_mem_pos_indexs[start : start + need_size]
        self.add_ref(rows)
        return rows, start, start + need_size

    @torch.no_grad()
    def alloc_kvcache_index(self, need_size: int) -> torch.Tensor | None:
        """The per-step entry point: ``need_size`` rows as int32, preferring consecutive ones.

        On the append-only path this is a view of a constant arange plus one in-place add on the
        device -- no host read, no synchronisation (a decode step calls it once)."""
        if self._bump_is_exact and self._bump_cursor + need_size <= self.max_num_tokens:
            start = self._bump_cursor
            self.kv_mem_use_state[start : start + need_size] += 1
            self._bump_cursor = start + need_size
            self.can_use_mem_size -= need_size
            return self.kv_mem_pos_indexs_int32[start : start + need_size]
        got = self.alloc_contiguous_kvcache(need_size)
        rows = got[0] if got is not None else self.alloc_kvcache(need_size)
        return None if rows is None else rows.to(torch.int32)

    # -------------------------------------------------------- reference counts #
    @torch.no_grad()
    def add_ref(self, token_index: torch.Tensor) -> None:
        """One more reference on each row; only rows that were free cost capacity."""
        token_index = token_index.long()
        newly_used = int((self.kv_mem_use_state[token_index] == 0).sum().item())
        self.can_use_mem_size -= newly_used
        self.kv_mem_use_state[token_index] += 1

    @torch.no_grad()
    def release_ref(self, token_index: torch.Tensor) -> None:
        """One reference less per occurrence in ``token_index`` (a row may be named several times: the engine
        releases prefill and every decode step in one concatenated tensor); rows reaching zero are free."""
        self._bump_is_exact = False            # holes: the cursor no longer describes the free list
        rows, counts = token_index.long().unique(return_counts=True)
        self.kv_mem_use_state[rows] -= counts.to(torch.int32)
        self.can_use_mem_size += int((self.kv_mem_use_state[rows] == 0).sum().item())

    @torch.no_grad()
    def free(self, free_index: torch.Tensor) -> None:
        self.release_ref(free_index.long())

    @torch.no_grad()
    def claim(self, num_rows: int) -> None:
        """Hand the first ``num_rows`` rows to an external owner (the slot table) permanently; the bump
        cursor resumes past them so both schemes can share the pool."""
        if num_rows > self.can_use_mem_size:
            raise ValueError(f"cannot claim {num_rows} rows: only {self.can_use_mem_size} are free")
        self.kv_mem_use_state[:num_rows] += 1
        self.can_use_mem_size -= num_rows
        self._bump_cursor = max(self._bump_cursor, num_rows)

    @torch.no_grad()
    def free_all(self) -> None:
        self.kv_mem_use_state.zero_()
        self.can_use_mem_size = self.max_num_tokens
        self._bump_cursor = 0
        self._bump_is_exact = True

"""Token-granular KV pool and its row allocator.

Keeps the layout the reference fixes as part of its executor surface
(lite_llama/executor/kv_cache_manager.py:136-372): one tensor per layer,
``[max_tokens, 2 * num_kv_heads, head_dim]`` with the K heads first, rows handed out by a bump
cursor (``alloc_kvcache_index``, kv_cache_manager.py:270-300) that ``free_all`` rewinds
(:364-372) and ``claim`` pre-reserves for the slot layout (:336-352). The reference's
reference-counted free list is only reached when requests are retired individually; the bump path
is what both the one-shot engine and the slot engine use on the hot path.
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
        self.can_use_mem_size = gpu_num_blocks
        self.kv_mem_pos_indexs_int32 = torch.arange(self.max_num_tokens, dtype=torch.int32, device=device)
        self._bump_cursor = 0
        self.gpu_kv_buffer = [
            torch.empty((self.max_num_tokens, 2 * num_kv_heads, head_dim), dtype=dtype, device=device)
            for _ in range(num_layers)
        ]

    def alloc_kvcache_index(self, need_size: int) -> torch.Tensor | None:
        """Next ``need_size`` rows as an int32 view (no device reads, no sync); ``None`` when short."""
        if self._bump_cursor + need_size > self.max_num_tokens:
            return None
        start = self._bump_cursor
        self._bump_cursor += need_size
        self.can_use_mem_size -= need_size
        return self.kv_mem_pos_indexs_int32[start : start + need_size]

    def claim(self, num_rows: int) -> None:
        """Hand the first ``num_rows`` rows to an external owner (the slot table) permanently."""
        if num_rows > self.max_num_tokens:
            raise ValueError(f"cannot claim {num_rows} rows from a pool of {self.max_num_tokens}")
        self._bump_cursor = max(self._bump_cursor, num_rows)
        self.can_use_mem_size = self.max_num_tokens - self._bump_cursor

    def free_all(self) -> None:
        self._bump_cursor = 0
        self.can_use_mem_size = self.max_num_tokens

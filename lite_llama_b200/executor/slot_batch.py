"""Slot layout of the KV pool + per-step attention metadata for continuous batching.

Same surface and meaning as the reference's ``SlotBatch`` (lite_llama/executor/slot_batch.py:28-220), which
``ContinuousBatchingEngine`` drives (engine/continuous_engine.py:157,298-419): slot ``s`` owns cache rows
``[s * max_seq_len, (s + 1) * max_seq_len)`` for good, so the request -> row table is the identity map written
once, reserving a request's cache is a host-side slot hand-out, and a steady-state decode step moves nothing
between host and device: lengths advance in place on the device and the write rows are a gather from the table.
The last slot is kept back as the home of the filler rows that pad an odd decode batch up to a captured CUDA-graph
width (ModelRunner.graph_batch_size).
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import TYPE_CHECKING

import torch

if TYPE_CHECKING:  # pragma: no cover
    from .model_runner import ModelRunner


class SlotBatch:
    def __init__(self, runner: "ModelRunner") -> None:
        self._runner = runner
        self._atten = runner.atten_info
        self.device = runner.device
        self.max_seq_len = runner.max_seq_len
        table = runner.b_req_tokens_table
        slots_total, width = table.shape
        # one slot is held back for graph-padding filler rows when there is one to spare
        self._filler_slot = slots_total - 1 if slots_total > 1 else None
        self.num_slots = slots_total - 1 if slots_total > 1 else 1
        table.copy_(torch.arange(slots_total * width, dtype=table.dtype, device=self.device).view(slots_total, width))
        if self._filler_slot is not None:
            lo = self._filler_slot * width
            for layer in runner.kv_cache_manager.gpu_kv_buffer:
                layer[lo : lo + width].zero_()      # filler rows attend here: keep it finite
        runner.kv_cache_manager.claim(slots_total * width)
        self._grid_rows = torch.arange(slots_total, dtype=torch.int32, device=self.device)
        self._dev_slots: torch.Tensor | None = None
        self._dev_lens: torch.Tensor | None = None
        self._cur_slots: list[int] = []
        self._cur_lens: list[int] = []

    # ------------------------------------------------------------------ steps #
    def begin_prefill(self, slots: Sequence[int], prompt_lens: Sequence[int]) -> None:
        """Attention metadata for a padded ``[n, max_prompt_len]`` prefill grid: token ``j`` of sequence ``i``
        lands in row ``j`` of slot ``slots[i]`` (padding positions write rows the sequence's own decode steps
        overwrite in order; attention never reads past ``b_seq_len``)."""
        longest = max(prompt_lens)
        if longest > self.max_seq_len:
            raise ValueError(f"prompt length {longest} exceeds max_seq_len {self.max_seq_len}")
        idx = self._upload(slots)
        at = self._atten
        at.b_req_idx = idx
        at.b_seq_len = self._upload(prompt_lens)
        at.max_actual_seq_len = longest
        at.cur_select_index = at.b_req_tokens_table[idx, :longest].reshape(-1)
        at.b_start_loc = self._grid_rows[: len(slots)] * longest
        self.reset()            # the running set changed: the next decode rebuilds its metadata

    def begin_decode(self, slots: Sequence[int], seq_lens: Sequence[int]) -> int:
        """Attention metadata for one decode step; ``seq_lens[i]`` is the length AFTER this step's token (its K/V
        goes to row ``seq_lens[i] - 1`` of the slot). Returns the submitted batch size, which exceeds
        ``len(slots)`` when the batch was padded to a captured graph width (discard the trailing logits)."""
        slots_p, lens_p = self._pad(list(slots), list(seq_lens))
        advanced = [n + 1 for n in self._cur_lens]
        if slots_p == self._cur_slots and lens_p == advanced:
            self._dev_lens += 1                      # same requests, one token further: no transfer
        else:
            self._dev_slots = self._upload(slots_p)
            self._dev_lens = self._upload(lens_p)
        self._cur_slots, self._cur_lens = slots_p, lens_p
        at = self._atten
        at.b_req_idx = self._dev_slots
        at.b_seq_len = self._dev_lens
        at.max_actual_seq_len = max(seq_lens)
        at.cur_select_index = at.b_req_tokens_table[self._dev_slots, self._dev_lens - 1]
        at.b_start_loc = None
        return len(slots_p)

    def reset(self) -> None:
        self._cur_slots, self._cur_lens = [], []

    @property
    def seq_lens(self) -> torch.Tensor:
        """Cache length per row of the batch last submitted (filler rows included)."""
        if self._dev_lens is None:
            raise RuntimeError("begin_decode() must run before seq_lens is read")
        return self._dev_lens

    # -------------------------------------------------------------- internals #
    def _pad(self, slots: list[int], seq_lens: list[int]) -> tuple[list[int], list[int]]:
        if self._filler_slot is None:
            return slots, seq_lens
        extra = self._runner.graph_batch_size(len(slots)) - len(slots)
        if extra <= 0:
            return slots, seq_lens
        # fillers carry the batch's own longest length so that every row advances by one per step
        fill_len = min(max(seq_lens), self.max_seq_len)
        return slots + [self._filler_slot] * extra, seq_lens + [fill_len] * extra

    def _upload(self, values: Sequence[int]) -> torch.Tensor:
        # a fresh tensor every time: the previous one may still be queued for the GPU
        return torch.tensor(list(values), dtype=torch.long, device=self.device)

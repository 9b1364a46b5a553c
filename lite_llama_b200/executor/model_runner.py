"""ModelRunner: owns the decoder, the KV pool and the per-step attention state.

Keeps the reference's surface (lite_llama/executor/model_runner.py:40-330): attributes
``kv_cache_manager``, ``b_req_tokens_table``, ``atten_info``, ``max_request_num``, ``max_seq_len``,
``vocab_size``, ``model``; methods ``prefill_alloc_kv_cache``, ``decode_alloc_kv_cache``,
``enable_cuda_graph``, ``graph_batch_size`` and ``forward``. ``build`` takes a geometry preset and
random-initialises the weights (no checkpoints exist offline) instead of a checkpoint directory.

Decode steps replay a captured CUDA graph per ``(batch, seq-bucket)`` exactly like the reference's
CUDAGraphManager (executor/cuda_graph.py:54-249): inputs are ``copy_``-ed into persistent buffers,
``max_actual_seq_len`` is baked at capture. Unlike the reference (model_runner.py:260-266) graphs
stay enabled under tensor parallelism: every rank captures the identical op sequence, so the
all-reduces inside the graph line up by construction.
"""

from __future__ import annotations

from typing import Any

import torch

from ..distributed import all_reduce_min, divide, get_tp_world_size
from ..kernels import argmax_rows, update_kv_index
from ..kernels._common import counters
from .attention_metadata import AttentionMetadata
from .decoder import Decoder, DecoderWeights, random_weights
from .kv_cache_manager import KVCacheManager
from .model_config import PRESETS, ModelConfig
from .slot_batch import SlotBatch

# executor/cuda_graph.py:27 stops at 128; 256 added so that a 256-wide continuous batch (BASELINE configs[3]) stays
# on the graph path
DEFAULT_BATCH_SIZES = (1, 2, 4, 8, 16, 32, 64, 128, 256)
DEFAULT_SEQ_LEN_BUCKETS = (256, 512, 1024, 2048, 4096)       # executor/cuda_graph.py:28


class ModelSpec:
    """What the engines ask about an architecture (reference: lite_llama/models/registry.py:24-40)."""

    def __init__(self, model_type: str, is_multimodal: bool = False) -> None:
        self.model_type = model_type
        self.is_multimodal = is_multimodal


class _CapturedDecode:
    """One captured decode step for a fixed batch size and seq-len bucket."""

    def __init__(self, runner: "ModelRunner", batch: int, bucket: int) -> None:
        dev = runner.device
        self.batch, self.bucket = batch, bucket
        self.input_ids = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.position_ids = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.b_req_idx = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.b_seq_len = torch.ones(batch, dtype=torch.int64, device=dev)
        self.cur_select_index = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.info = AttentionMetadata(
            kv_buffer=runner.kv_cache_manager.gpu_kv_buffer,
            cur_select_index=self.cur_select_index,
            b_req_tokens_table=runner.b_req_tokens_table,
            b_req_idx=self.b_req_idx,
            b_seq_len=self.b_seq_len,
            max_actual_seq_len=bucket,
        )
        self.graph = torch.cuda.CUDAGraph()
        # warm up on a side stream (lazy module loads, workspace allocation), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                runner.model.forward(self.input_ids, self.position_ids, self.info)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.logits = runner.model.forward(self.input_ids, self.position_ids, self.info)

    def replay(self, input_ids, position_ids, info: AttentionMetadata) -> torch.Tensor:
        n = input_ids.shape[0]
        self.input_ids[:n].copy_(input_ids)
        self.position_ids[:n].copy_(position_ids)
        self.b_req_idx[:n].copy_(info.b_req_idx)
        self.b_seq_len[:n].copy_(info.b_seq_len)
        self.cur_select_index[:n].copy_(info.cur_select_index)
        self.graph.replay()
        return self.logits[:n]


class GreedyDecodeLoop:
    """A whole greedy decode step captured as ONE CUDA graph: token-table update, model forward,
    argmax, and the on-device bookkeeping the reference's hot loop does per step
    (lite_llama/engine/llm_engine.py:137-213: forward -> decode_alloc_kv_cache -> sample).

    Replaying it advances every sequence by one token with zero host work, which is what the
    device-resident throughput figure in bench.py measures. The KV rows of step i are
    ``first_row + i*batch + arange(batch)`` -- exactly what the bump allocator hands out
    (kv_cache_manager.py:286-292) -- so the graph carries its own cursor.
    """

    def __init__(self, runner: "ModelRunner", batch: int, bucket: int) -> None:
        dev = runner.device
        self.runner, self.batch, self.bucket = runner, batch, bucket
        self.input_ids = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.position_ids = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.b_req_idx = torch.arange(batch, dtype=torch.int32, device=dev)
        self.b_seq_len = torch.ones(batch, dtype=torch.int32, device=dev)
        self.cur_select_index = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.next_tokens = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.info = AttentionMetadata(
            kv_buffer=runner.kv_cache_manager.gpu_kv_buffer, cur_select_index=self.cur_select_index,
            b_req_tokens_table=runner.b_req_tokens_table, b_req_idx=self.b_req_idx, b_seq_len=self.b_seq_len,
            max_actual_seq_len=bucket)
        self.graph: torch.cuda.CUDAGraph | None = None
        self.launches_per_step = 0
        self._steps_left = 0          # host-side bound: rows reserved for the loop / room left in the bucket
        self._reserved: tuple[int, int] | None = None      # (first row, steps) handed out by begin()

    def _step_ops(self) -> None:
        # state on entry: input_ids = last sampled token, b_seq_len = history length INCLUDING it,
        # position = b_seq_len - 1, cur_select_index = the rows reserved for it
        update_kv_index(self.info.b_req_tokens_table, self.b_req_idx, self.b_seq_len, self.cur_select_index)
        logits = self.runner.model.forward(self.input_ids, self.position_ids, self.info)
        argmax_rows(logits[:, -1, :], out=self.next_tokens)            # greedy sampling without torch's 45 us reduce
        self.input_ids.copy_(self.next_tokens.unsqueeze(1))
        self.position_ids.add_(1)
        self.b_seq_len.add_(1)
        self.cur_select_index.add_(self.batch)

    def begin(self, last_tokens: torch.Tensor, seq_lens: torch.Tensor, b_req_idx: torch.Tensor,
              first_row: int | None = None, max_steps: int | None = None) -> None:
        """Seed the loop after prefill: ``last_tokens`` are the first sampled tokens, ``seq_lens`` the
        prompt lengths (the history length becomes ``seq_lens + 1`` once the token is appended).

        The loop advances its KV cursor on the device, so its rows are reserved HERE, through the allocator:
        ``max_steps`` steps need ``batch * max_steps`` consecutive rows (``first_row=None``); ``step()`` refuses to
        run past them or past the graph's seq-len bucket. Passing ``first_row`` re-seeds a loop on rows reserved
        by an earlier ``begin`` (the stress test replays one reservation many times)."""
        longest = int(seq_lens.max().item()) + 1
        room = self.bucket - longest + 1               # steps until a history would exceed the bucket
        if first_row is None:
            if max_steps is None:
                raise ValueError("begin() needs max_steps to reserve the loop's KV rows")
            rows = self.runner.kv_cache_manager.alloc_kvcache_index(self.batch * max_steps)
            if rows is None:
                raise RuntimeError(f"KV cache exhausted: the loop needs {self.batch * max_steps} rows")
            first_row = int(rows[0].item())
            if int(rows[-1].item()) - first_row != rows.numel() - 1:
                raise RuntimeError("the loop needs consecutive KV rows (allocate it before any partial free)")
            self._reserved = (first_row, max_steps)
        elif self._reserved is None or self._reserved[0] != first_row:
            self._reserved = (first_row, max_steps if max_steps is not None else room)   # caller-managed rows
        self._steps_left = max(0, min(self._reserved[1], room))
        self.input_ids.copy_(last_tokens.view(-1, 1))
        self.b_seq_len.copy_((seq_lens + 1).to(torch.int32))
        self.position_ids.copy_(seq_lens.view(-1, 1).to(torch.int64))
        self.b_req_idx.copy_(b_req_idx.to(torch.int32))
        self.cur_select_index.copy_(torch.arange(first_row, first_row + self.batch, dtype=torch.int32,
                                                 device=self.runner.device))

    def capture(self) -> None:
        from .. import _lib

        lib = _lib.load()
        snapshot = [t.clone() for t in (self.input_ids, self.position_ids, self.b_seq_len, self.cur_select_index)]
        side = torch.cuda.Stream(device=self.runner.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step_ops()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((self.input_ids, self.position_ids, self.b_seq_len, self.cur_select_index), snapshot):
            dst.copy_(src)
        before = lib.b2l_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_ops()
        self.launches_per_step = int(lib.b2l_launch_count() - before)
        # capture does not execute: the state tensors still hold the snapshot

    def step(self) -> None:
        if self._steps_left <= 0:
            raise RuntimeError("GreedyDecodeLoop.step(): past the KV rows reserved by begin() or past the seq-len bucket "
                               f"({self.bucket}); nothing was launched")
        self._steps_left -= 1
        self.graph.replay()


class ModelRunner:
    def __init__(self, config: ModelConfig, weights: DecoderWeights, max_gpu_num_blocks: int,
                 device: str = "cuda", dtype: torch.dtype = torch.float16) -> None:
        self.config = config
        self.device = device
        self.dtype = dtype
        self.num_layers = config.num_layers
        self.num_kv_heads = divide(config.num_kv_heads, get_tp_world_size(), "key/value heads")
        self.head_dim = config.head_dim
        self.vocab_size = config.vocab_size
        self.max_seq_len = config.max_seq_len
        # every rank must agree on the pool size (model_runner.py:84-89)
        max_gpu_num_blocks = all_reduce_min(max_gpu_num_blocks, device)
        self.kv_cache_manager = KVCacheManager(self.num_layers, self.num_kv_heads, self.head_dim,
                                               max_gpu_num_blocks, dtype=dtype, device=device)
        self.max_request_num = max(1, max_gpu_num_blocks // self.max_seq_len)
        self.b_req_tokens_table = torch.zeros((self.max_request_num, self.max_seq_len), dtype=torch.int32,
                                              device=device)
        self.atten_info = AttentionMetadata()
        self.atten_info.kv_buffer = self.kv_cache_manager.gpu_kv_buffer
        self.atten_info.b_req_tokens_table = self.b_req_tokens_table
        self.model = Decoder(config, weights, max_pos=self.max_seq_len, device=device, dtype=dtype)
        self.spec = ModelSpec(config.name)
        self._slot_batch: SlotBatch | None = None
        self._graphs: dict[tuple[int, int], _CapturedDecode] = {}
        self._graph_batches: tuple[int, ...] = ()
        self._graph_buckets: tuple[int, ...] = ()
        counters(torch.device(device))  # allocate the split-K ticket buffer outside any capture

    # ------------------------------------------------------------------ build #
    @classmethod
    def build(cls, model: str | ModelConfig, max_seq_len: int, max_gpu_num_blocks: int | None = None,
              device: str = "cuda", quantization: str | None = None, seed: int = 0,
              dtype: torch.dtype = torch.float16, num_layers: int | None = None,
              draw_device: str | None = None) -> "ModelRunner":
        """``model`` is a preset name (see ``PRESETS``) or a ModelConfig; ``quantization`` one of
        ``None | "int4" | "int8" | "fp8" | "smoothquant"`` (the reference's ``--quantization`` values,
        models/quantization/config.py:39-54)."""
        cfg = PRESETS[model] if isinstance(model, str) else model
        cfg = cfg.with_(max_seq_len=max_seq_len)
        if num_layers is not None:
            cfg = cfg.with_(num_layers=num_layers)
        weights = random_weights(cfg, quantization or "fp16", device=device, seed=seed, dtype=dtype,
                                 draw_device=draw_device)
        if max_gpu_num_blocks is None:
            max_gpu_num_blocks = 32 * max_seq_len
        return cls(cfg, weights, max_gpu_num_blocks, device=device, dtype=dtype)

    # --------------------------------------------------------- kv allocation #
    def _init_req_tokens_table(self, b_req_idx, b_seq_len, alloc_index, max_prompt_len) -> torch.Tensor:
        """Padded-rectangle layout: sequence i's j-th token is row ``i*max_prompt_len + j`` of the
        flattened batch (model_runner.py:153-179). Vectorised, no host loop."""
        batch = b_seq_len.shape[0]
        grid = alloc_index.view(batch, max_prompt_len)
        cols = torch.arange(max_prompt_len, device=self.device)
        valid = cols[None, :] < b_seq_len[:, None]
        rows = b_req_idx.long()[:, None].expand(batch, max_prompt_len)
        self.b_req_tokens_table[rows[valid], cols[None, :].expand(batch, max_prompt_len)[valid]] = grid[valid]
        return (torch.arange(batch, device=self.device, dtype=torch.int32) * max_prompt_len)

    def prefill_alloc_kv_cache(self, max_prompt_len, actual_prompt_lens, b_req_idx) -> torch.Tensor:
        batch_size = len(actual_prompt_lens)
        self.atten_info.b_req_idx = b_req_idx
        self.atten_info.cur_select_index = self.kv_cache_manager.alloc_kvcache_index(max_prompt_len * batch_size)
        if self.atten_info.cur_select_index is None:
            raise RuntimeError("KV cache exhausted during prefill allocation")
        self.atten_info.b_seq_len = actual_prompt_lens
        self.atten_info.max_actual_seq_len = max_prompt_len
        self.atten_info.b_start_loc = self._init_req_tokens_table(
            b_req_idx, actual_prompt_lens, self.atten_info.cur_select_index, max_prompt_len)
        return self.atten_info.cur_select_index

    def decode_alloc_kv_cache(self, batch_size) -> torch.Tensor:
        """One row per sequence for the next step; ``b_seq_len`` is bumped BEFORE update_kv_index
        writes position ``b_seq_len - 1`` (model_runner.py:200-218)."""
        self.atten_info.cur_select_index = self.kv_cache_manager.alloc_kvcache_index(batch_size)
        if self.atten_info.cur_select_index is None:
            raise RuntimeError("KV cache exhausted during decode allocation")
        self.atten_info.b_seq_len += 1
        self.atten_info.max_actual_seq_len += 1
        update_kv_index(self.atten_info.b_req_tokens_table, self.atten_info.b_req_idx,
                        self.atten_info.b_seq_len, self.atten_info.cur_select_index)
        return self.atten_info.cur_select_index

    # ------------------------------------------------------------- inference #
    def enable_cuda_graph(self, batch_sizes: tuple[int, ...] = DEFAULT_BATCH_SIZES,
                          seq_len_buckets: tuple[int, ...] = DEFAULT_SEQ_LEN_BUCKETS) -> None:
        """Capture one decode graph per ``(batch, seq bucket)`` (model_runner.py:243-308). Capture
        OOM falls back to eager, as in the reference (:301-307)."""
        seq_len_buckets = tuple(b for b in seq_len_buckets if b <= self.max_seq_len)
        batch_sizes = tuple(b for b in batch_sizes if b <= self.max_request_num)
        if not seq_len_buckets or not batch_sizes:
            return
        if self.kv_cache_manager.can_use_mem_size != self.kv_cache_manager.max_num_tokens:
            # the capture warm-up writes dummy K/V through row 0 of request 0 (like the reference's,
            # executor/cuda_graph.py:100-119): with live rows that would clobber a request's first token
            raise RuntimeError("enable_cuda_graph() must run before any KV rows are allocated (its warm-up writes row 0)")
        try:
            for b in batch_sizes:
                for s in seq_len_buckets:
                    if (b, s) not in self._graphs:
                        self._graphs[(b, s)] = _CapturedDecode(self, b, s)
        except torch.cuda.OutOfMemoryError:
            self._graphs.clear()
            return
        self._graph_batches = tuple(sorted({k[0] for k in self._graphs}))
        self._graph_buckets = tuple(sorted({k[1] for k in self._graphs}))

    def enable_slot_kv_cache(self) -> SlotBatch:
        """Switch the KV pool to the fixed-slot layout continuous batching uses (model_runner.py:221-234).
        Idempotent; mutually exclusive with the one-shot ``*_alloc_kv_cache`` path (the slot table claims every
        row it names). Call it after :meth:`enable_cuda_graph` so that batch padding sees the captured widths."""
        if self._slot_batch is None:
            self._slot_batch = SlotBatch(self)
        return self._slot_batch

    def graph_batch_size(self, batch_size: int) -> int:
        for b in self._graph_batches:
            if b >= batch_size:
                return b
        return batch_size

    def _try_replay(self, input_ids, position_ids) -> torch.Tensor | None:
        if not self._graphs or input_ids.shape[1] != 1:
            return None
        batch = input_ids.shape[0]
        b = next((x for x in self._graph_batches if x >= batch), None)
        s = next((x for x in self._graph_buckets if x >= self.atten_info.max_actual_seq_len), None)
        if b is None or s is None or (b, s) not in self._graphs:
            return None
        if b != batch:
            return None  # the engine pads the batch itself (graph_batch_size); never guess filler rows here
        return self._graphs[(b, s)].replay(input_ids, position_ids, self.atten_info)

    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor,
                multi_modal_inputs: dict[str, Any] | None = None,
                last_token_index: torch.Tensor | None = None) -> torch.Tensor:
        """One model step: captured graph for eligible decode steps, eager otherwise
        (model_runner.py:310-330)."""
        replayed = self._try_replay(input_ids, position_ids)
        if replayed is not None:
            return replayed
        return self.model.forward(input_ids, position_ids, self.atten_info, last_token_index)

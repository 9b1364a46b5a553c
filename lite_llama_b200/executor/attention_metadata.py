"""Per-step KV-cache bookkeeping handed to the attention ops.

Same fields, meaning and dtypes as the reference dataclass
(lite_llama/executor/attention_metadata.py:19-44): the runner mutates them in place before each
forward, and the decode CUDA graph reads them from persistent buffers.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch


@dataclass
class AttentionMetadata:
    kv_buffer: list[torch.Tensor] = field(default_factory=list)   # per layer [max_tokens, 2*Hkv, D]
    cur_select_index: torch.Tensor | None = None                  # int32 rows written this step
    b_req_tokens_table: torch.Tensor | None = None                # int32 [max_requests, max_seq_len]
    b_start_loc: torch.Tensor | None = None                       # prefill: start of each sequence
    b_req_idx: torch.Tensor | None = None                         # slot of each batch row
    b_seq_len: torch.Tensor | None = None                         # history length incl. this step
    max_actual_seq_len: int = 0                                   # host int, sizes the split grid

"""KVCacheManager and SlotBatch: the reference's own test list for the allocator
(/root/reference/tests/executor/test_kv_cache_manager.py:47-215: initial state, scattered / contiguous allocation,
refusal without side effects, reference counts, bump fast path and when it must switch off) restated against this
package's implementation, plus the slot layout the continuous-batching engine drives
(lite_llama/executor/slot_batch.py:56-220). Runs on CPU tensors (pure host logic + torch indexing)."""

from __future__ import annotations

import pytest
import torch

from lite_llama_b200.executor.attention_metadata import AttentionMetadata
from lite_llama_b200.executor.kv_cache_manager import KVCacheManager
from lite_llama_b200.executor.slot_batch import SlotBatch

N = 9


@pytest.fixture
def m() -> KVCacheManager:
    return KVCacheManager(num_layers=2, num_kv_heads=4, head_dim=64, gpu_num_blocks=N, dtype=torch.float32, device="cpu")


def test_initial_state_and_buffers(m):
    assert m.can_use_mem_size == N and m.kv_mem_use_state.numel() == N and not m.kv_mem_use_state.any()
    assert len(m.gpu_kv_buffer) == 2 and m.gpu_kv_buffer[0].shape == (N, 8, 64)
    assert m.gpu_kv_buffer[0].data_ptr() != m.gpu_kv_buffer[1].data_ptr()


def test_scattered_alloc_and_total_refusal(m):
    idx = m.alloc_kvcache(3)
    assert idx.numel() == 3 and (m.kv_mem_use_state[idx] == 1).all() and m.can_use_mem_size == N - 3
    assert m.alloc_kvcache(N) is None
    assert m.can_use_mem_size == N - 3 and int(m.kv_mem_use_state.sum()) == 3      # nothing changed


def test_contiguous_alloc_is_a_real_run_search(m):
    idx, lo, hi = m.alloc_contiguous_kvcache(4)
    assert idx.tolist() == list(range(lo, hi)) and hi - lo == 4 and m.can_use_mem_size == N - 4
    m.free_all()
    m.alloc_kvcache(5)                       # rows 0-4
    m.kv_mem_use_state[7] = 1                # hole at 7 leaves runs of 2 and 1
    m.can_use_mem_size -= 1
    assert m.alloc_contiguous_kvcache(3) is None
    assert m.alloc_contiguous_kvcache(2)[0].tolist() == [5, 6]


def test_contiguous_alloc_reuses_released_rows(m):
    idx, _, _ = m.alloc_contiguous_kvcache(4)
    m.release_ref(idx[:2])
    again = m.alloc_contiguous_kvcache(2)
    assert again is not None and again[0].numel() == 2 and m.can_use_mem_size == N - 4


def test_reference_counts(m):
    idx = m.alloc_kvcache(2)
    m.add_ref(idx)
    assert (m.kv_mem_use_state[idx] == 2).all() and m.can_use_mem_size == N - 2
    m.release_ref(idx)
    assert (m.kv_mem_use_state[idx] == 1).all() and m.can_use_mem_size == N - 2
    m.release_ref(idx)
    assert not m.kv_mem_use_state[idx].any() and m.can_use_mem_size == N
    # one tensor naming a row several times (the engine releases prefill + every decode step at once)
    a = m.alloc_kvcache_index(3)
    m.add_ref(a[:1].long())
    m.release_ref(torch.cat([a.long(), a[:1].long()]))
    assert not m.kv_mem_use_state.any() and m.can_use_mem_size == N
    m.free(m.alloc_kvcache(2))
    assert m.can_use_mem_size == N


def test_bump_fast_path(m):
    a, b = m.alloc_kvcache_index(3), m.alloc_kvcache_index(2)
    assert a.tolist() == [0, 1, 2] and b.tolist() == [3, 4] and a.dtype == torch.int32
    assert (m.kv_mem_use_state[a.long()] == 1).all() and m.can_use_mem_size == N - 5


def test_partial_free_switches_the_bump_path_off_until_free_all(m):
    first = m.alloc_kvcache_index(4)
    m.release_ref(first[:2].long())
    assert m._bump_is_exact is False
    assert sorted(m.alloc_kvcache_index(2).tolist()) == [0, 1]           # the holes are reused, not rows 4-5
    m.free_all()
    assert m._bump_is_exact is True and m._bump_cursor == 0
    assert m.alloc_kvcache_index(2).tolist() == [0, 1]


def test_bump_falls_back_at_the_end_and_cycles_do_not_leak(m):
    m.alloc_kvcache_index(N)
    m.free_all()
    m.alloc_kvcache_index(N - 2)
    assert m.alloc_kvcache_index(2).numel() == 2
    assert m.alloc_kvcache_index(1) is None
    for _ in range(10):
        m.free_all()
        assert m.alloc_kvcache_index(4).numel() == 4
    m.free_all()
    assert m.can_use_mem_size == N and not m.kv_mem_use_state.any()


def test_claim_shares_the_pool_with_the_bump_cursor(m):
    m.claim(4)
    assert m.can_use_mem_size == N - 4 and (m.kv_mem_use_state[:4] == 1).all()
    assert m.alloc_kvcache_index(2).tolist() == [4, 5]
    with pytest.raises(ValueError):
        m.claim(N)


class _Runner:
    """The slice of ModelRunner a SlotBatch touches."""

    def __init__(self, slots=4, width=8, graph_sizes=()):
        self.device = "cpu"
        self.max_seq_len = width
        self.kv_cache_manager = KVCacheManager(1, 1, 8, slots * width, dtype=torch.float32, device="cpu")
        self.b_req_tokens_table = torch.zeros((slots, width), dtype=torch.int32)
        self.atten_info = AttentionMetadata()
        self.atten_info.b_req_tokens_table = self.b_req_tokens_table
        self._sizes = tuple(graph_sizes)

    def graph_batch_size(self, n):
        return next((s for s in self._sizes if s >= n), n)


def test_slot_batch_layout_and_claim():
    r = _Runner()
    sb = SlotBatch(r)
    assert sb.num_slots == 3                                             # the last slot backs filler rows
    assert r.b_req_tokens_table.tolist() == [list(range(s * 8, s * 8 + 8)) for s in range(4)]
    assert r.kv_cache_manager.can_use_mem_size == 0                      # every row now belongs to a slot
    single = SlotBatch(_Runner(slots=1))
    assert single.num_slots == 1


def test_slot_batch_prefill_metadata():
    r = _Runner()
    sb = SlotBatch(r)
    sb.begin_prefill([2, 0], [3, 5])
    at = r.atten_info
    assert at.b_req_idx.tolist() == [2, 0] and at.b_seq_len.tolist() == [3, 5] and at.max_actual_seq_len == 5
    assert at.cur_select_index.tolist() == list(range(16, 21)) + list(range(0, 5))   # padded [n, 5] grid, row-major
    assert at.b_start_loc.tolist() == [0, 5]
    with pytest.raises(ValueError):
        sb.begin_prefill([1], [9])


def test_slot_batch_decode_steady_state_moves_nothing():
    r = _Runner()
    sb = SlotBatch(r)
    sb.begin_prefill([1, 2], [3, 4])
    assert sb.begin_decode([1, 2], [4, 5]) == 2
    at = r.atten_info
    assert at.cur_select_index.tolist() == [8 + 3, 16 + 4] and at.max_actual_seq_len == 5 and at.b_start_loc is None
    lens_tensor = at.b_seq_len
    sb.begin_decode([1, 2], [5, 6])
    assert at.b_seq_len is lens_tensor and lens_tensor.tolist() == [5, 6]            # advanced in place
    assert at.cur_select_index.tolist() == [8 + 4, 16 + 5]
    sb.begin_decode([2], [7])                                                        # a request left: rebuilt
    assert at.b_seq_len is not lens_tensor and at.b_req_idx.tolist() == [2] and sb.seq_lens.tolist() == [7]


def test_slot_batch_pads_to_a_captured_graph_width():
    r = _Runner(slots=6, graph_sizes=(1, 2, 4))
    sb = SlotBatch(r)
    n = sb.begin_decode([0, 1, 2], [3, 5, 4])
    at = r.atten_info
    assert n == 4 and at.b_req_idx.tolist() == [0, 1, 2, 5] and at.b_seq_len.tolist() == [3, 5, 4, 5]
    assert at.max_actual_seq_len == 5
    # fillers advance with the batch, so the in-place fast path survives padding
    t = at.b_seq_len
    sb.begin_decode([0, 1, 2], [4, 6, 5])
    assert at.b_seq_len is t and t.tolist() == [4, 6, 5, 6]
    sb.reset()
    sb.begin_decode([0, 1, 2], [5, 7, 6])
    assert at.b_seq_len is not t

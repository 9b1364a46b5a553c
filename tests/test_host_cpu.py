"""Host-side logic that needs no GPU: KV row allocator, padded-rectangle token table, the quantisers
against the reference's packers (golden), the TP shard arithmetic and a world_size-2 gloo all-reduce."""

from __future__ import annotations

import os
import subprocess
import sys
import textwrap

import pytest
import torch

from b2l_testutil import ROOT, load_golden
from lite_llama_b200.executor.kv_cache_manager import KVCacheManager
from lite_llama_b200.executor.model_config import PRESETS
from lite_llama_b200.quantization import params


def test_kv_manager_bump_allocator_and_claim():
    m = KVCacheManager(num_layers=2, num_kv_heads=2, head_dim=8, gpu_num_blocks=32, device="cpu")
    assert len(m.gpu_kv_buffer) == 2 and m.gpu_kv_buffer[0].shape == (32, 4, 8)
    a = m.alloc_kvcache_index(10)
    b = m.alloc_kvcache_index(4)
    assert a.dtype == torch.int32 and a.tolist() == list(range(10)) and b.tolist() == [10, 11, 12, 13]
    assert m.can_use_mem_size == 18
    assert m.alloc_kvcache_index(19) is None          # short: the reference also returns None
    m.free_all()
    assert m.alloc_kvcache_index(3).tolist() == [0, 1, 2]
    m.free_all()
    m.claim(16)
    assert m.alloc_kvcache_index(2).tolist() == [16, 17]
    with pytest.raises(ValueError):
        m.claim(100)


def test_prefill_table_uses_the_padded_rectangle_layout():
    # sequence i's j-th token is flattened row i*max_prompt_len + j (model_runner.py:153-179 in the reference)
    from lite_llama_b200.executor.model_runner import ModelRunner

    r = ModelRunner.__new__(ModelRunner)
    r.device = "cpu"
    r.b_req_tokens_table = torch.zeros(4, 16, dtype=torch.int32)
    alloc = torch.arange(100, 100 + 3 * 5, dtype=torch.int32)
    start = ModelRunner._init_req_tokens_table(r, torch.tensor([2, 0, 3]), torch.tensor([5, 2, 4]), alloc, 5)
    assert start.tolist() == [0, 5, 10]
    assert r.b_req_tokens_table[2, :5].tolist() == [100, 101, 102, 103, 104]
    assert r.b_req_tokens_table[0, :3].tolist() == [105, 106, 0]
    assert r.b_req_tokens_table[3, :5].tolist() == [110, 111, 112, 113, 0]
    assert r.b_req_tokens_table[1].abs().sum() == 0


def test_int4_packer_is_bit_identical_to_the_reference_packer():
    for c in load_golden("w4a16")["cases"]:
        q, s, z = params.quantize_int4_groupwise(c["w"], c["group_size"])
        assert torch.equal(q, c["qweight"]) and torch.equal(s, c["scales"]) and torch.equal(z, c["zeros"])


def test_int8_and_fp8_packers_match_the_reference_packers():
    for c in load_golden("w8a16")["cases"]:
        q, s = params.quantize_int8_per_channel(c["w"])
        assert torch.equal(q, c["int8_channel"]["qweight"]) and torch.equal(s, c["int8_channel"]["scales"])
        q, s = params.quantize_fp8_per_channel(c["w"])
        assert torch.equal(q, c["fp8_channel"]["qweight"]) and torch.equal(s, c["fp8_channel"]["scales"])


def test_presets_have_the_public_geometries():
    l3 = PRESETS["llama-3-8b"]
    assert (l3.hidden_size, l3.num_layers, l3.num_heads, l3.num_kv_heads, l3.head_dim, l3.intermediate_size,
            l3.vocab_size) == (4096, 32, 32, 8, 128, 14336, 128256)
    q7 = PRESETS["qwen2.5-7b"]
    assert (q7.hidden_size, q7.num_heads, q7.num_kv_heads, q7.intermediate_size, q7.qkv_bias) == (3584, 28, 4, 18944, True)
    for cfg in PRESETS.values():
        cfg.validate()


_TP_SCRIPT = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, {root!r})
    from lite_llama_b200.distributed import init_tensor_parallel, all_reduce_tp, all_reduce_min, get_tp_rank, get_tp_world_size, divide, destroy_tensor_parallel
    init_tensor_parallel(backend="gloo")
    r, w = get_tp_rank(), get_tp_world_size()
    assert w == 2
    # row-parallel partial sums: each rank holds x[:, shard] @ W[:, shard].T, the all-reduce restores x @ W.T
    torch.manual_seed(0)
    x, W = torch.randn(3, 8), torch.randn(5, 8)
    k = divide(8, w, "K")
    part = x[:, r*k:(r+1)*k] @ W[:, r*k:(r+1)*k].T
    full = all_reduce_tp(part.clone())
    torch.testing.assert_close(full, x @ W.T, rtol=1e-5, atol=1e-5)
    assert all_reduce_min(100 - 10 * r) == 90
    try:
        divide(7, 2, "heads"); raise SystemExit("divide must reject 7/2")
    except ValueError:
        pass
    destroy_tensor_parallel()
    print("rank", r, "ok")
""")


def test_tensor_parallel_all_reduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "tp2.py"
    script.write_text(_TP_SCRIPT.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_stream_k_partition_arithmetic():
    """The (tile, raw block) split of the quantised GEMM (csrc/gemm_quant.cu: WorkRange / sk_owner) restated in
    integers: CTA ranges tile the space exactly, `sk_owner` inverts the range formula, pieces add up, and the
    host's `workspace slots per tile` (b2l_quant_gemm_plan) is the largest number of CTAs sharing a tile."""
    import ctypes
    import random

    from lite_llama_b200 import _lib

    def owner(lin, ctas, total):
        return ((lin + 1) * ctas - 1) // total

    def check(n_tiles, nrb_tile, ctas):
        total = n_tiles * nrb_tile
        cover = [0] * total
        for c in range(ctas):
            g0, g1 = c * total // ctas, (c + 1) * total // ctas
            assert g1 > g0
            for lin in range(g0, g1):
                cover[lin] += 1
                assert owner(lin, ctas, total) == c
            t0 = g0 // nrb_tile
            npieces = (g1 - 1) // nrb_tile - t0 + 1
            blocks = 0
            for i in range(npieces):
                base = (t0 + i) * nrb_tile
                lo, hi = max(g0 - base, 0), min(g1 - base, nrb_tile)
                assert 0 <= lo < hi <= nrb_tile
                blocks += hi - lo
            assert blocks == g1 - g0
            # what the flag-in-data hand-over of shared tiles relies on (csrc/gemm_quant.cu, epilogue + end of kernel):
            # at most two pieces of a range are shared (its first and its last tile), and a CTA that is slot 0 of a
            # shared tile -- the tile's reducer -- meets that tile as the LAST piece of its range (so its weight ring is
            # idle when it parks its own partial sum there, and it reduces at most one tile)
            shared = []
            for i in range(npieces):
                t = t0 + i
                c_lo, c_hi = owner(t * nrb_tile, ctas, total), owner((t + 1) * nrb_tile - 1, ctas, total)
                assert c_lo <= c <= c_hi
                if c_hi > c_lo:
                    shared.append(i)
                    if c == c_lo:
                        assert i == npieces - 1, "the reducer's tile must be the last piece of its range"
            assert len(shared) <= 2 and all(i in (0, npieces - 1) for i in shared)
        assert all(v == 1 for v in cover)
        return max(owner((t + 1) * nrb_tile - 1, ctas, total) - owner(t * nrb_tile, ctas, total) + 1
                   for t in range(n_tiles))

    lib = _lib.load()
    for n, k in [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (640, 1792), (1000, 2304)]:
        out = (ctypes.c_int * 10)()
        _lib.check(lib.b2l_quant_gemm_plan(0, 32, n, k, 128, out, None))
        sk_ctas, slots, nrb_tile, n_tiles = out[0], out[1], out[2], out[3]
        assert sk_ctas > 0 and nrb_tile == k // 256 and n_tiles == (n + 127) // 128
        assert check(n_tiles, nrb_tile, sk_ctas) == slots
    rng = random.Random(0)
    for _ in range(300):
        nt, nrb = rng.randint(1, 300), rng.randint(1, 64)
        check(nt, nrb, max(1, min(nt * nrb // 2, 148)))


def test_fp8_block_quantiser_and_sub_block_tensor_parallel_shards():
    """Fine-grained FP8 (128 x 128 block scales, the Qwen3-30B-A3B-FP8 checkpoint format): the product's quantiser
    equals the oracle's; refining the scale grid (reblock_scales) leaves every dequantised value unchanged; and a
    tensor-parallel shard that starts inside a 128-block (768 / 4 = 192 columns per rank) dequantises to the same
    values as the matching slice of the full matrix -- the case the reference refuses to load
    (lite_llama/models/linear.py:164-177)."""
    from lite_llama_b200.quantization.params import quantize_fp8_blockwise, reblock_scales
    from oracle import ops

    torch.manual_seed(0)
    w = torch.randn(3, 256, 768) * 0.02
    q, s = quantize_fp8_blockwise(w)
    q2, s2 = ops.quantize_fp8_blockwise(w)
    assert torch.equal(q, q2) and torch.equal(s, s2) and s.shape == (3, 2, 6)
    full = ops.dequant_w8(q[1], s[1], 128, 128)
    assert float((full - w[1]).abs().max()) < 0.02 * 4 / 8          # e4m3: 3 mantissa bits
    s64 = reblock_scales(reblock_scales(s, 128, 64, 1), 128, 64, 2)
    assert torch.equal(ops.dequant_w8(q[1], s64[1], 64, 64), full)
    for rank in range(4):                                            # contracted-dim shards of 192 columns
        lo, hi = rank * 192, (rank + 1) * 192
        shard = ops.dequant_w8(q[1][:, lo:hi].contiguous(), s64[1][:, lo // 64:hi // 64].contiguous(), 64, 64)
        assert torch.equal(shard, full[:, lo:hi])
    with pytest.raises(ValueError):
        reblock_scales(s, 128, 48, 1)


def test_moe_expert_sharding_matches_the_reference_partition():
    """make_moe_weights: rank r keeps intermediate columns [r*I/tp, (r+1)*I/tp) of EVERY expert -- gate rows and up
    rows of w1, contracted columns of w2 (lite_llama/models/weights.py:129-130) -- so the per-rank partial outputs
    sum to the unsharded block's output (checked with the oracle's fused_moe in fp32-friendly fp16)."""
    from lite_llama_b200.executor.decoder import make_moe_weights
    from oracle import ops

    cfg = PRESETS["tiny-qwen3-moe"]
    g = torch.Generator().manual_seed(1)
    e, im, h = cfg.num_experts, cfg.moe_intermediate_size, cfg.hidden_size
    router, gate, up, down = (torch.randn(*sh, generator=g) * 0.05 for sh in ((e, h), (e, im, h), (e, im, h), (e, h, im)))
    x = (torch.randn(5, h, generator=g) * 0.5).half()
    full = make_moe_weights(cfg, router, gate, up, down, "fp16", 1, 0, torch.float16)
    w, ids = ops.route_topk(x, full.router, cfg.num_experts_per_tok, True)
    want = ops.fused_moe(x, full.w1, full.w2, w, ids).float()
    parts = []
    for r in range(2):
        m = make_moe_weights(cfg, router, gate, up, down, "fp16", 2, r, torch.float16)
        assert m.w1.shape == (e, im, h) and m.w2.shape == (e, h, im // 2)
        parts.append(ops.fused_moe(x, m.w1, m.w2, w, ids).float())
    torch.testing.assert_close(parts[0] + parts[1], want, rtol=2e-2, atol=2e-3)
    m8 = make_moe_weights(cfg, router, gate, up, down, "fp8", 4, 1, torch.float16)     # 128 / 4 = 32 columns per rank
    assert m8.group_n == 32 and m8.group_k == 32 and m8.w1.dtype == torch.uint8
    assert m8.w1_scale.shape == (e, 2, h // 32) and m8.w2_scale.shape == (e, h // 32, 1)


def test_awq_and_gptq_checkpoint_tensors_convert_to_the_reference_layout():
    """AutoAWQ GEMM / AutoGPTQ tensors -> reference W4A16 layout (quantization/awq.py). The checkpoint side is packed
    here with each format's published rule (AutoAWQ: eight output channels per word, order [0,2,4,6,1,3,5,7];
    AutoGPTQ: eight input channels per word, zero points stored minus one), the converted tensors must dequantise
    (oracle.ops.dequant_int4) to exactly (q - z) * s of the source integers."""
    from lite_llama_b200.quantization.awq import AWQ_ORDER, awq_to_reference, gptq_to_reference
    from oracle import ops

    g = torch.Generator().manual_seed(0)
    K_, N_, gs = 256, 64, 128
    q = torch.randint(0, 16, (K_, N_), generator=g)                     # [K, N] integer weights
    z = torch.randint(0, 16, (K_ // gs, N_), generator=g)
    s = (torch.rand(K_ // gs, N_, generator=g) * 0.02 + 0.001).half()
    want = ((q - z.repeat_interleave(gs, 0)).float() * s.float().repeat_interleave(gs, 0)).t()     # [N, K]

    def pack_cols(t, order):                                            # eight columns per int32 word
        w = torch.zeros(t.shape[0], t.shape[1] // 8, dtype=torch.int64)
        for p, o in enumerate(order):
            w |= t[:, o::8].to(torch.int64) << (4 * p)
        return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)

    qw, sc, zp = awq_to_reference(pack_cols(q, AWQ_ORDER), pack_cols(z, AWQ_ORDER), s)
    assert qw.shape == (N_, K_ // 8) and sc.shape == zp.shape == (N_, K_ // gs)
    assert torch.equal(ops.dequant_int4(qw, sc, zp, gs), want)

    rows = torch.zeros(K_ // 8, N_, dtype=torch.int64)                  # GPTQ: eight input rows per word
    for i in range(8):
        rows |= q[i::8].to(torch.int64) << (4 * i)
    rows = torch.where(rows >= 2**31, rows - 2**32, rows).to(torch.int32)
    qw2, sc2, zp2 = gptq_to_reference(rows, pack_cols((z - 1) % 16, range(8)), s, g_idx=torch.arange(K_) // gs)
    # zero points of 0 wrap to 15 in the 4-bit "minus one" storage; restrict the comparison to the others
    ok = (z.t() > 0).repeat_interleave(gs, 1)
    got = ops.dequant_int4(qw2, sc2, zp2, gs)
    assert torch.equal(got[ok], want[ok]) and torch.equal(qw2, qw)
    with pytest.raises(ValueError):
        gptq_to_reference(rows, pack_cols((z - 1) % 16, range(8)), s, g_idx=torch.randperm(K_) // gs)

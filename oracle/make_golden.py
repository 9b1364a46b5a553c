"""Generate ``tests/golden/*.pt`` from the REFERENCE ITSELF (run in the build container only).

Two sources, both imported from ``/root/reference`` (read-only, never copied):
  * the reference's own Triton kernels, executed on CPU tensors through Triton's interpreter
    (``TRITON_INTERPRET=1``), i.e. the actual ``lite_llama.kernels`` code path;
  * its CPU-runnable helpers: the quantisers (models/quantization/params/*.py),
    ``moe_align_block_size`` (kernels/fused_moe.py:45-99) and ``tests/reference.py``.

The fixtures are small (a few hundred KiB in total), seeded, and committed together with this
script; ``tests/test_oracle_golden.py`` pins ``oracle/ops.py`` against them on CPU and the
``-m gpu`` tests compare the CUDA kernels with the same files on the B200.

Usage:  python oracle/make_golden.py      (needs /root/reference; ~1 min)
"""

from __future__ import annotations

import math
import os
import sys
from pathlib import Path

os.environ["TRITON_INTERPRET"] = "1"
REF = "/root/reference"
sys.path.insert(0, REF)

import torch  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
LOG2E = 1.4426950408889634


def _save(name: str, obj: dict) -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(obj, OUT / f"{name}.pt")
    size = (OUT / f"{name}.pt").stat().st_size
    print(f"  {name}.pt  {size / 1024:.1f} KiB")


def main() -> None:
    from lite_llama.kernels import (
        flash_attention2_no_pad,
        flash_decoding,
        fused_moe,
        moe_align_block_size,
        rope_emb_forward,
        skip_rmsnorm,
        swiglu_forward,
        update_kv_buffer,
        update_kv_index,
    )
    from lite_llama.kernels.quantization import smoothquant_matmul, w4a16_matmul, w8a16_matmul
    from lite_llama.models.quantization import quantize_int4_groupwise, quantize_int8_per_channel
    from lite_llama.models.quantization.params.fp8 import quantize_fp8_per_channel
    from tests import reference as ref

    f16 = torch.float16
    g = torch.Generator().manual_seed(0)

    def randn(*shape, scale=1.0, dtype=f16):
        return (torch.randn(*shape, generator=g) * scale).to(dtype)

    # ---- skip_rmsnorm ---------------------------------------------------------------
    cases = []
    for rows, n, with_res in [(5, 64, True), (3, 896, True), (4, 128, False), (2, 4096, True)]:
        x, w = randn(rows, n), randn(n)
        r = randn(rows, n) if with_res else None
        r_in = None if r is None else r.clone()
        y, r_out = skip_rmsnorm(x.clone(), r_in, w, 1e-5)
        cases.append({"x": x, "residual": r, "weight": w, "eps": 1e-5, "y": y, "residual_out": r_out if with_res else None})
    _save("skip_rmsnorm", {"cases": cases})

    # ---- swiglu ----------------------------------------------------------------------
    cases = []
    for rows, n in [(3, 128), (2, 4864)]:
        a, b = randn(rows, n, scale=2.0), randn(rows, n)
        cases.append({"a": a, "b": b, "out": swiglu_forward(a, b)})
    _save("swiglu", {"cases": cases})

    # ---- rope ------------------------------------------------------------------------
    cases = []
    for bsz, sl, hq, hk, d in [(2, 3, 4, 2, 64), (1, 5, 14, 2, 128), (3, 1, 8, 8, 32)]:
        q, k = randn(bsz * sl, hq, d), randn(bsz * sl, hk, d)
        ang = torch.rand(bsz, sl, d // 2, generator=g) * 6.28
        cos = torch.cat([ang.cos(), ang.cos()], -1).to(f16)
        sin = torch.cat([ang.sin(), ang.sin()], -1).to(f16)
        qo, ko = rope_emb_forward(q.clone(), k.clone(), cos, sin, bsz, sl)
        cases.append({"q": q, "k": k, "cos": cos, "sin": sin, "batch": bsz, "seq": sl, "q_out": qo, "k_out": ko,
                      "q_ref_fp32": ref.rope_half_split(q, cos, sin, bsz, sl)})
    _save("rope", {"cases": cases})

    # ---- kv cache ops ----------------------------------------------------------------
    vals = randn(6, 4, 32)
    sel = torch.tensor([9, 2, 17, 4, 0, 11], dtype=torch.int32)
    buf = randn(20, 4, 32)
    buf_out = buf.clone()
    update_kv_buffer(vals, sel, buf_out)
    table = torch.zeros(4, 16, dtype=torch.int32)
    req = torch.tensor([2, 0, 3], dtype=torch.int32)
    lens = torch.tensor([5, 16, 1], dtype=torch.int32)
    selidx = torch.tensor([100, 200, 300], dtype=torch.int32)
    table_out = table.clone()
    update_kv_index(table_out, req, lens, selidx)
    _save("kv_ops", {"values": vals, "select": sel, "buffer": buf, "buffer_out": buf_out, "table": table,
                     "b_req_idx": req, "b_seq_len": lens, "select_index": selidx, "table_out": table_out})

    # ---- flash_decoding --------------------------------------------------------------
    cases = []
    for seq_lens, hq, hk, d, scattered, req_idx in [
        ([1, 17], 4, 4, 64, False, None),
        ([33, 130], 8, 2, 64, True, None),
        ([40, 150, 64], 14, 2, 128, False, [2, 0, 1]),
        ([129], 8, 1, 32, True, None),
    ]:
        max_tokens = 512
        q = randn(len(seq_lens), hq, d, scale=0.3)
        kc, vc = randn(max_tokens, hk, d), randn(max_tokens, hk, d)
        width = max(seq_lens)
        table = torch.zeros(len(seq_lens), width, dtype=torch.int32)
        if scattered:
            perm = torch.randperm(max_tokens, generator=g).to(torch.int32)
            off = 0
            for i, n in enumerate(seq_lens):
                table[i, :n] = perm[off : off + n]
                off += n
        else:
            for i, n in enumerate(seq_lens):
                table[i, :n] = torch.arange(i * width, i * width + n, dtype=torch.int32)
        b_seq = torch.tensor(seq_lens, dtype=torch.int32)
        b_req = torch.arange(len(seq_lens), dtype=torch.int32) if req_idx is None else torch.tensor(req_idx, dtype=torch.int32)
        if req_idx is not None:
            # lengths belong to batch rows; make every slot long enough for whichever row reads it
            full = torch.zeros(len(seq_lens), width, dtype=torch.int32)
            for i in range(len(seq_lens)):
                full[i] = torch.arange(i * width, (i + 1) * width, dtype=torch.int32)
            table = full
        scale = 1.0 / math.sqrt(d)
        out = flash_decoding(q, kc, vc, scale, table, b_req, b_seq, width)
        ref_out = ref.paged_decode_attention(q, kc, vc, table[b_req.long()], b_seq, scale)
        cases.append({"q": q, "k_cache": kc, "v_cache": vc, "table": table, "b_req_idx": b_req, "b_seq_len": b_seq,
                      "max_len": width, "scale": scale, "out": out, "ref_fp32": ref_out})
    _save("flash_decoding", {"cases": cases})

    # ---- flash_attention2_no_pad -----------------------------------------------------
    cases = []
    for seq_lens, hq, hk, d in [([16], 4, 4, 64), ([7, 100, 33], 8, 2, 64), ([1, 64], 14, 2, 128), ([40, 24], 4, 2, 32)]:
        total = sum(seq_lens)
        q, k, v = randn(total, hq, d, scale=0.3), randn(total, hk, d, scale=0.3), randn(total, hk, d, scale=0.3)
        starts = torch.tensor([sum(seq_lens[:i]) for i in range(len(seq_lens))], dtype=torch.int32)
        lens = torch.tensor(seq_lens, dtype=torch.int32)
        scale = 1.0 / math.sqrt(d)
        out = flash_attention2_no_pad(q, k, v, scale * LOG2E, starts, lens, max(seq_lens))
        cases.append({"q": q, "k": k, "v": v, "b_start_loc": starts, "b_seq_len": lens, "max_len": max(seq_lens),
                      "scale": scale, "out": out, "ref_fp32": ref.varlen_causal_attention(q, k, v, starts, lens, scale)})
    _save("flash_attention_nopad", {"cases": cases})

    # ---- quantisers + quantised GEMMs ------------------------------------------------
    cases = []
    for m, n, k, gs in [(1, 64, 256, 32), (5, 96, 256, 128), (33, 130, 384, 128)]:
        x = randn(m, k, scale=0.5)
        w = randn(n, k, scale=0.05, dtype=torch.float32)
        qw, sc, zp = quantize_int4_groupwise(w, gs)
        bias = randn(n, dtype=torch.float32) if m == 5 else None
        cases.append({"x": x, "w": w, "qweight": qw, "scales": sc, "zeros": zp, "group_size": gs, "bias": bias,
                      "out": w4a16_matmul(x, qw, sc, zp, group_size=gs, bias=bias)})
    _save("w4a16", {"cases": cases})

    cases = []
    for m, n, k in [(1, 128, 256), (8, 256, 384), (33, 130, 384)]:
        x = randn(m, k, scale=0.5)
        w = randn(n, k, scale=0.05, dtype=torch.float32)
        qw_fp8 = w.to(torch.float8_e4m3fn).view(torch.uint8)
        sc_blk = torch.rand((n + 127) // 128, (k + 127) // 128, generator=g) + 0.5
        qw_i8, sc_i8 = quantize_int8_per_channel(w)
        qw_f8c, sc_f8c = quantize_fp8_per_channel(w)
        cases.append({
            "x": x, "w": w,
            "fp8_block": {"qweight": qw_fp8, "scales": sc_blk, "group_n": 128, "group_k": 128,
                          "out": w8a16_matmul(x, qw_fp8, sc_blk, group_n=128, group_k=128)},
            "int8_channel": {"qweight": qw_i8, "scales": sc_i8, "group_n": 1, "group_k": k,
                             "out": w8a16_matmul(x, qw_i8, sc_i8, group_n=1, group_k=k)},
            "fp8_channel": {"qweight": qw_f8c, "scales": sc_f8c, "group_n": 1, "group_k": k,
                            "out": w8a16_matmul(x, qw_f8c, sc_f8c, group_n=1, group_k=k)},
        })
    _save("w8a16", {"cases": cases})

    cases = []
    for m, n, k in [(1, 64, 256), (8, 128, 512), (33, 130, 384)]:
        x = randn(m, k, scale=0.5)
        w = randn(n, k, scale=0.05, dtype=torch.float32)
        scale = w.abs().amax(-1) / 127.0
        qw = (w / scale.unsqueeze(-1)).round().clamp(-127, 127).to(torch.int8)
        bias = randn(n, dtype=torch.float32) if m == 8 else None
        cases.append({"x": x, "qweight": qw, "weight_scales": scale, "bias": bias,
                      "out": smoothquant_matmul(x, qw, scale, bias=bias)})
    _save("w8a8", {"cases": cases})

    # ---- moe -------------------------------------------------------------------------
    cases = []
    for ids, blk, ne in [([[2, 0], [1, 2]], 4, 3), ([[5, 5]] * 4, 4, 8)]:
        t = torch.tensor(ids, dtype=torch.int32)
        s, e, npost = moe_align_block_size(t, blk, ne)
        cases.append({"topk_ids": t, "block_size": blk, "num_experts": ne, "sorted_ids": s, "expert_ids": e, "num_post": npost})
    t = torch.randint(0, 16, (9, 4), generator=g, dtype=torch.int32)
    s, e, npost = moe_align_block_size(t, 16, 16)
    cases.append({"topk_ids": t, "block_size": 16, "num_experts": 16, "sorted_ids": s, "expert_ids": e, "num_post": npost})
    _save("moe_align", {"cases": cases})

    cases = []
    for tokens, ne, tk, hidden, inter in [(3, 8, 2, 128, 64), (9, 16, 4, 128, 64)]:
        hs = randn(tokens, hidden, scale=hidden**-0.5)
        w1 = randn(ne, 2 * inter, hidden, scale=hidden**-0.5)
        w2 = randn(ne, hidden, inter, scale=inter**-0.5)
        ids = torch.randint(0, ne, (tokens, tk), generator=g)
        tw = torch.softmax(torch.randn(tokens, tk, generator=g), dim=-1).to(f16)
        cases.append({"hidden": hs, "w1": w1, "w2": w2, "topk_weights": tw, "topk_ids": ids,
                      "out": fused_moe(hs, w1, w2, tw, ids)})
    # fp8 block-quantised experts (config 5 layout)
    tokens, ne, tk, hidden, inter = 5, 4, 2, 256, 128
    hs = randn(tokens, hidden, scale=hidden**-0.5)
    w1 = randn(ne, 2 * inter, hidden, scale=hidden**-0.5, dtype=torch.float32).to(torch.float8_e4m3fn).view(torch.uint8)
    w2 = randn(ne, hidden, inter, scale=inter**-0.5, dtype=torch.float32).to(torch.float8_e4m3fn).view(torch.uint8)
    s1 = torch.rand(ne, (2 * inter + 127) // 128, (hidden + 127) // 128, generator=g) + 0.5
    s2 = torch.rand(ne, (hidden + 127) // 128, (inter + 127) // 128, generator=g) + 0.5
    ids = torch.randint(0, ne, (tokens, tk), generator=g)
    tw = torch.softmax(torch.randn(tokens, tk, generator=g), dim=-1).to(f16)
    cases.append({"hidden": hs, "w1": w1, "w2": w2, "topk_weights": tw, "topk_ids": ids, "w1_scale": s1, "w2_scale": s2,
                  "group_n": 128, "group_k": 128,
                  "out": fused_moe(hs, w1, w2, tw, ids, w1_scale=s1, w2_scale=s2, group_n=128, group_k=128)})
    _save("fused_moe", {"cases": cases})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not mounted; golden vectors can only be regenerated in the build container")
    torch.manual_seed(0)
    main()

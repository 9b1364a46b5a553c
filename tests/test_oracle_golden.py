"""Pin the CPU oracle (oracle/ops.py) against the golden vectors recorded FROM THE REFERENCE.

``tests/golden/*.pt`` were produced by ``oracle/make_golden.py``: the reference's own Triton kernels
run through Triton's CPU interpreter, plus its CPU-runnable quantisers / align helper. If these
pass, the oracle the GPU tests rely on states the same function as the reference.

Tolerances: integer/byte/index results are compared exactly; fp16 results within 2 fp16 ulps of the
reference output (the interpreter rounds every fp16 op separately, the oracle computes in fp32).
"""

from __future__ import annotations

import pytest
import torch

from b2l_testutil import load_golden
from oracle import ops

FP16_TOL = dict(rtol=2e-3, atol=1e-3)


def _close(actual, expected, **tol):
    torch.testing.assert_close(actual.float(), expected.float(), **(tol or FP16_TOL))


def test_skip_rmsnorm_matches_reference_kernel():
    for c in load_golden("skip_rmsnorm")["cases"]:
        y, r = ops.skip_rmsnorm(c["x"], c["residual"], c["weight"], c["eps"])
        _close(y, c["y"])
        if c["residual"] is not None:
            assert torch.equal(r, c["residual_out"]), "residual sum must round identically"


def test_swiglu_matches_reference_kernel():
    for c in load_golden("swiglu")["cases"]:
        _close(ops.swiglu(c["a"], c["b"]), c["out"])


def test_rope_matches_reference_kernel():
    # the kernel rotates in fp16 (rope_emb.py:64-83): allow 2 ulps at the data's magnitude (|x| < 8)
    for c in load_golden("rope")["cases"]:
        for name in ("q", "k"):
            got = ops.rope_half_split(c[name], c["cos"], c["sin"], c["batch"], c["seq"])
            _close(got, c[f"{name}_out"], rtol=2e-3, atol=4e-3)
        # and the reference's own fp32 restatement is reproduced exactly
        assert torch.equal(ops.rope_half_split(c["q"], c["cos"], c["sin"], c["batch"], c["seq"]), c["q_ref_fp32"])


def test_kv_cache_ops_match_reference_kernels():
    g = load_golden("kv_ops")
    buf = ops.update_kv_buffer(g["values"], g["select"], g["buffer"].clone())
    assert torch.equal(buf, g["buffer_out"])
    table = ops.update_kv_index(g["table"].clone(), g["b_req_idx"], g["b_seq_len"], g["select_index"])
    assert torch.equal(table, g["table_out"])


def test_flash_decoding_matches_reference_kernel():
    for c in load_golden("flash_decoding")["cases"]:
        got = ops.paged_decode_attention(
            c["q"], c["k_cache"], c["v_cache"], c["table"], c["b_req_idx"], c["b_seq_len"], c["scale"]
        )
        _close(got, c["out"])
        _close(got, c["ref_fp32"])


def test_flash_attention_nopad_matches_reference_kernel():
    for c in load_golden("flash_attention_nopad")["cases"]:
        got = ops.varlen_causal_attention(c["q"], c["k"], c["v"], c["b_start_loc"], c["b_seq_len"], c["scale"])
        assert torch.equal(got, c["ref_fp32"])  # same definition as tests/reference.py
        # rows of real tokens only: the kernel leaves padding rows uninitialised (NaN in the interpreter)
        valid = torch.zeros(c["q"].shape[0], dtype=torch.bool)
        for s, n in zip(c["b_start_loc"].tolist(), c["b_seq_len"].tolist()):
            valid[s : s + n] = True
        _close(got[valid], c["out"][valid])


def test_int4_packer_and_w4a16_match_reference():
    for c in load_golden("w4a16")["cases"]:
        qw, s, z = ops.quantize_int4_groupwise(c["w"], c["group_size"])
        assert torch.equal(qw, c["qweight"]) and torch.equal(s, c["scales"]) and torch.equal(z, c["zeros"])
        nib = ops.unpack_int4(c["qweight"])
        assert nib.min() >= 0 and nib.max() <= 15
        # bit-exact unpack order: repacking the nibbles reproduces the words
        shifts = torch.arange(8, dtype=torch.int64) * 4
        repacked = (nib.reshape(nib.shape[0], -1, 8) << shifts).sum(-1)
        assert torch.equal(repacked.to(torch.int32), c["qweight"])
        got = ops.w4a16_matmul(c["x"], c["qweight"], c["scales"], c["zeros"], c["group_size"], c["bias"])
        _close(got, c["out"])


def test_w8a16_matches_reference_kernel():
    for c in load_golden("w8a16")["cases"]:
        for key in ("fp8_block", "int8_channel", "fp8_channel"):
            d = c[key]
            got = ops.w8a16_matmul(c["x"], d["qweight"], d["scales"], d["group_n"], d["group_k"])
            _close(got, d["out"])


def test_fp8_bit_trick_is_exact_for_all_finite_codes():
    codes = torch.arange(256, dtype=torch.int32).to(torch.uint8)
    finite = (codes & 0x7F) != 0x7F
    via_trick = ops.fp8_e4m3_bits_to_float(codes)[finite]
    via_cast = codes.view(torch.float8_e4m3fn).float()[finite]
    assert torch.equal(via_trick, via_cast)


def test_smoothquant_matches_reference_kernel_exactly():
    # truncation-toward-zero activation codes + exact int32 accumulation: bit-identical fp16 output
    for c in load_golden("w8a8")["cases"]:
        got = ops.smoothquant_matmul(c["x"], c["qweight"], c["weight_scales"], c["bias"])
        _close(got, c["out"], rtol=1e-3, atol=5e-4)


def test_moe_align_matches_reference():
    for c in load_golden("moe_align")["cases"]:
        s, e, n = ops.moe_align_block_size(c["topk_ids"], c["block_size"], c["num_experts"])
        assert torch.equal(s, c["sorted_ids"])
        assert torch.equal(e, c["expert_ids"])
        assert torch.equal(n, c["num_post"])


def test_moe_align_known_answer():
    # tests/kernels/test_fused_moe.py:42-56 in the reference
    s, e, n = ops.moe_align_block_size(torch.tensor([[2, 0], [1, 2]], dtype=torch.int32), 4, 3)
    assert int(n) == 12
    assert [v for v in s[:12].tolist() if v != 4] == [1, 2, 0, 3]
    assert e[:3].tolist() == [0, 1, 2]


def test_fused_moe_matches_reference_kernel():
    for c in load_golden("fused_moe")["cases"]:
        got = ops.fused_moe(
            c["hidden"], c["w1"], c["w2"], c["topk_weights"], c["topk_ids"],
            c.get("w1_scale"), c.get("w2_scale"), c.get("group_n", 0), c.get("group_k", 0),
        )
        _close(got, c["out"], rtol=2e-3, atol=1e-4)


@pytest.mark.parametrize("residual", [True, False])
def test_known_answers_norm(residual):
    # tests/kernels/test_norm_activation.py:53-67: 2 + 3 = 5 everywhere -> unit RMS -> y == weight
    x = torch.full((2, 64), 2.0, dtype=torch.float16)
    r = torch.full((2, 64), 3.0, dtype=torch.float16) if residual else None
    w = torch.arange(64, dtype=torch.float16) / 64
    y, res = ops.skip_rmsnorm(x, r, w, 1e-6)
    torch.testing.assert_close(y.float(), w.float().expand(2, 64), rtol=1e-3, atol=1e-3)
    assert torch.equal(res, torch.full((2, 64), 5.0, dtype=torch.float16) if residual else x)


def test_w4_pack_is_a_permutation_of_the_reference_layout():
    # the kernel-native repack must carry exactly the reference's nibbles (w4a16.py:9-12) and table entries
    torch.manual_seed(0)
    qw = torch.randint(-(2**31), 2**31 - 1, (130, 24), dtype=torch.int64).to(torch.int32)
    packed = ops.pack_w4_qweight(qw)
    assert packed.dtype == torch.int32 and packed.shape == qw.shape
    assert torch.equal(ops.unpack_w4_packed(packed), ops.unpack_int4(qw))
    # known answer: nibbles 0..7 in K order -> positions hold K indices [0,2,4,6,1,3,5,7]
    one = torch.tensor([[0x76543210]], dtype=torch.int32)
    assert int(ops.pack_w4_qweight(one)[0, 0]) == 0x75316420
    tab = torch.randn(130, 6)
    pt = ops.pack_w4_table(tab)
    assert pt.shape == (2, 6, 128)
    for n in (0, 1, 127, 128, 129):
        assert torch.equal(pt[n // 128, :, n % 128], tab[n])
    assert float(pt[1, :, 2:].abs().max()) == 0.0

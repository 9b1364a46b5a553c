"""CPU oracle: plain-torch restatement of the reference's hot-path operators.

TEST INFRASTRUCTURE ONLY. Nothing under ``lite_llama_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do,
and only as the checker (or the timed CPU baseline), never as the product path.

Every function restates one reference operator and cites the file:line it follows (paths relative
to /root/reference). Parity is PINNED: ``tests/golden/`` holds outputs of the reference's own
Triton kernels (run here through Triton's CPU interpreter by ``oracle/make_golden.py``) and of its
CPU-runnable helpers (quantisers, ``moe_align_block_size``, ``tests/reference.py``); the oracle is
checked against every one of them in ``tests/test_oracle_golden.py``.

All functions run on CPU tensors, loop where the reference definition loops, and compute in fp32
unless the reference computes in the storage dtype (RoPE).
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# attention (tests/reference.py:21-126)
# --------------------------------------------------------------------------- #
def _expand_kv_heads(x: torch.Tensor, num_groups: int) -> torch.Tensor:
    """kv_head = q_head // groups, i.e. repeat_interleave on the head axis (tests/reference.py:21-34)."""
    return x.repeat_interleave(num_groups, dim=0)


def varlen_causal_attention(q, k, v, b_start_loc, b_seq_len, sm_scale):
    """Prefill oracle (tests/reference.py:37-81; kernel: kernels/flashattention2_nopad.py:45-169).

    ``sm_scale`` is the plain ``1/sqrt(D)`` (no log2e). Rows outside every sequence stay zero.
    """
    groups = q.shape[1] // k.shape[1]
    out = torch.zeros_like(q)
    for i in range(b_seq_len.shape[0]):
        start, length = int(b_start_loc[i]), int(b_seq_len[i])
        sl = slice(start, start + length)
        qi = q[sl].transpose(0, 1).float()
        ki = _expand_kv_heads(k[sl].transpose(0, 1).float(), groups)
        vi = _expand_kv_heads(v[sl].transpose(0, 1).float(), groups)
        scores = (qi @ ki.transpose(-1, -2)) * sm_scale
        causal = torch.ones(length, length, dtype=torch.bool).tril()
        scores = scores.masked_fill(~causal, float("-inf"))
        out[sl] = (scores.softmax(dim=-1) @ vi).transpose(0, 1).to(out.dtype)
    return out


def paged_decode_attention(q, k_cache, v_cache, b_req_tokens_table, b_req_idx, b_seq_len, sm_scale):
    """Decode oracle (tests/reference.py:84-126) with the slot indirection of the kernel
    (kernels/flashdecoding.py:66-69): row ``i`` reads ``table[b_req_idx[i], :b_seq_len[i]]``."""
    groups = q.shape[1] // k_cache.shape[1]
    out = torch.zeros_like(q)
    for i in range(q.shape[0]):
        length = int(b_seq_len[i])
        rows = b_req_tokens_table[int(b_req_idx[i]), :length].long()
        ki = _expand_kv_heads(k_cache[rows].transpose(0, 1).float(), groups)
        vi = _expand_kv_heads(v_cache[rows].transpose(0, 1).float(), groups)
        qi = q[i].float().unsqueeze(1)
        probs = ((qi @ ki.transpose(-1, -2)) * sm_scale).softmax(dim=-1)
        out[i] = (probs @ vi).squeeze(1).to(out.dtype)
    return out


# --------------------------------------------------------------------------- #
# norm / activation / rope (tests/reference.py:129-196)
# --------------------------------------------------------------------------- #
def rmsnorm(x, weight, eps=1e-5):
    """fp32 reduce, round to the activation dtype BEFORE the weight multiply
    (tests/reference.py:129-133; kernels/skip_rmsnorm.py:186-189)."""
    xf = x.float()
    scale = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * scale).to(x.dtype) * weight


def skip_rmsnorm(x, residual, weight, eps=1e-5):
    """``(normed, new_residual)`` (tests/reference.py:136-152; kernels/skip_rmsnorm.py:157-189).

    The kernel normalises the *unrounded* fp32 sum and stores the rounded sum back as the residual."""
    if residual is None:
        return rmsnorm(x, weight, eps), x
    summed_f32 = x.float() + residual.float()
    scale = torch.rsqrt(summed_f32.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (summed_f32 * scale).to(x.dtype) * weight, summed_f32.to(x.dtype)


def swiglu(gate, up):
    """``silu(gate_f32) * up`` (tests/reference.py:155-161; kernels/swiglu.py:36-41)."""
    return (F.silu(gate.float()) * up.float()).to(gate.dtype)


def rope_half_split(x, cos, sin, batch_size, seq_len):
    """Half-split RoPE reading the first D/2 of cos/sin (tests/reference.py:164-196;
    kernels/rope_emb.py:49-83). fp32 restatement; the kernel works in the storage dtype."""
    head_dim = x.shape[-1]
    half = head_dim // 2
    c = cos[..., :half].reshape(batch_size * seq_len, 1, half).float()
    s = sin[..., :half].reshape(batch_size * seq_len, 1, half).float()
    x1 = x[..., :half].float()
    x2 = x[..., half : 2 * half].float()
    rotated = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    if 2 * half < head_dim:
        rotated = torch.cat([rotated, x[..., 2 * half :].float()], dim=-1)
    return rotated.to(x.dtype)


# --------------------------------------------------------------------------- #
# KV-cache bookkeeping (kernels/update_kv_buffer.py:15-89, update_kv_index.py:15-88)
# --------------------------------------------------------------------------- #
def update_kv_buffer(kv_values, select_index, kv_buffer):
    """``kv_buffer[select_index[i]] = kv_values[i]`` in place; returns the buffer."""
    kv_buffer[select_index.long()] = kv_values
    return kv_buffer


def update_kv_index(table, b_req_idx, b_seq_len, select_index):
    """``table[b_req_idx[i], b_seq_len[i]-1] = select_index[i]`` in place; returns the table."""
    table[b_req_idx.long(), b_seq_len.long() - 1] = select_index.to(table.dtype)
    return table


# --------------------------------------------------------------------------- #
# quantised weights: packers (golden encoders) and dequantisers
# --------------------------------------------------------------------------- #
def quantize_int4_groupwise(weight, group_size=128):
    """Restates models/quantization/params/int4.py:12-49 (the only int4 producer in the reference)."""
    n, k = weight.shape
    if k % group_size != 0:
        raise ValueError(f"in_features {k} must be a multiple of group_size {group_size}")
    w = weight.float().reshape(n, k // group_size, group_size)
    w_min, w_max = w.amin(dim=-1), w.amax(dim=-1)
    scale = (w_max - w_min).clamp(min=1e-5) / 14.0
    zero = (-w_min / scale).round().clamp(0, 15)
    q = (w / scale.unsqueeze(-1) + zero.unsqueeze(-1)).round().clamp(0, 15).to(torch.int32)
    q = q.reshape(n, -1, 8)
    shifts = torch.arange(8, dtype=torch.int32) * 4
    packed = (q << shifts[None, None, :]).sum(dim=-1)
    return packed.to(torch.int32), scale.float(), zero.float()


def unpack_int4(qweight):
    """``[N, K/8]`` int32 -> ``[N, K]`` integer nibbles, nibble j of word i = K index 8i+j
    (kernels/quantization/w4a16.py:9-12,98-101; tests/kernels/test_quantization.py:84-91)."""
    word = qweight.to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(8, dtype=torch.int64) * 4
    nib = (word.unsqueeze(-1) >> shifts) & 0xF
    return nib.reshape(qweight.shape[0], -1)


def dequant_int4(qweight, scales, zeros, group_size):
    """``(nibble - zero) * scale`` per group, fp32 (kernels/quantization/w4a16.py:101-104)."""
    nib = unpack_int4(qweight).float()
    s = scales.float().repeat_interleave(group_size, dim=1)
    z = zeros.float().repeat_interleave(group_size, dim=1)
    return (nib - z) * s


def w4a16_matmul(x, qweight, scales, zeros, group_size=128, bias=None):
    """fp32 accumulate of x @ dequant(W).T (+ bias), rounded to x.dtype
    (kernels/quantization/w4a16.py:82-127; tests/kernels/test_quantization.py:84-103)."""
    out = x.float() @ dequant_int4(qweight, scales, zeros, group_size).T
    if bias is not None:
        out = out + bias.float()
    return out.to(x.dtype)


W4_PACK_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)   # K index (within a word) held by nibble position p of the packed word


def pack_w4_qweight(qweight):
    """Kernel-native nibble order of ``include/b200_lite_kernels.h: b2l_w4a16_pack_qweight``: a pure permutation
    of the reference word (kernels/quantization/w4a16.py:9-12: nibble j = K index 8i+j) -- position p of the
    packed word holds the reference nibble ``W4_PACK_ORDER[p]``."""
    word = qweight.to(torch.int64) & 0xFFFFFFFF
    out = torch.zeros_like(word)
    for pos, kk in enumerate(W4_PACK_ORDER):
        out |= ((word >> (4 * kk)) & 0xF) << (4 * pos)
    out = torch.where(out >= 2**31, out - 2**32, out)
    return out.to(torch.int32)


def unpack_w4_packed(qweight_packed):
    """Inverse view of :func:`pack_w4_qweight`: ``[N, K]`` nibbles in K order from the packed words."""
    word = qweight_packed.to(torch.int64) & 0xFFFFFFFF
    n = word.shape[0]
    nib = torch.zeros(n, word.shape[1], 8, dtype=torch.int64)
    for pos, kk in enumerate(W4_PACK_ORDER):
        nib[:, :, kk] = (word >> (4 * pos)) & 0xF
    return nib.reshape(n, -1)


def pack_w4_table(table):
    """``[N, G]`` fp32 scales or zero points -> ``[ceil(N/128), G, 128]`` (b2l_w4a16_pack_table): tile-major,
    weight row fastest, rows past N zero."""
    n, g = table.shape
    tiles = (n + 127) // 128
    padded = torch.zeros(tiles * 128, g, dtype=torch.float32)
    padded[:n] = table.float()
    return padded.view(tiles, 128, g).permute(0, 2, 1).contiguous()


def quantize_int8_per_channel(weight):
    """models/quantization/params/int8.py:12-28."""
    scale = weight.abs().amax(dim=-1, keepdim=True).float() / 127.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    qweight = (weight.float() / scale).round().clamp_(-127, 127).to(torch.int8)
    return qweight, scale


def quantize_fp8_per_channel(weight):
    """models/quantization/params/fp8.py:16-30 (uint8 container holding e4m3 bits)."""
    scale = weight.abs().amax(dim=-1, keepdim=True).float() / 448.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    q = (weight.float() / scale).clamp_(-448.0, 448.0).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), scale


def quantize_fp8_blockwise(weight, block=128):
    """Fine-grained FP8 checkpoint format (lite_llama/models/quantization/config.py:33 FP8_BLOCK = 128): one fp32
    scale per ``block x block`` tile = tile absmax / 448, codes = round-to-nearest e4m3 of ``w / scale``;
    dequantisation is ``code * scale`` per tile (lite_llama/executor/weight_utils.py:59-72)."""
    *lead, n, k = weight.shape
    nb, kb = (n + block - 1) // block, (k + block - 1) // block
    w = weight.float()
    if nb * block != n or kb * block != k:
        w = F.pad(w, (0, kb * block - k, 0, nb * block - n))
    tiles = w.reshape(*lead, nb, block, kb, block)
    amax = tiles.abs().amax(dim=(-3, -1))
    scales = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    q = (tiles / scales[..., :, None, :, None]).reshape(*lead, nb * block, kb * block)[..., :n, :k]
    return q.to(torch.float8_e4m3fn).view(torch.uint8).contiguous(), scales.contiguous()


def fp8_e4m3_bits_to_float(q_u8):
    """The reference's bit trick (kernels/quantization/w8a16.py:49-62) followed by its x256
    (w8a16.py:120): ``((b&0x80)<<8)|((b&0x7F)<<7)`` viewed as fp16, times 256. Bit-exact for all
    254 finite e4m3 codes; the two NaN codes become +-480 like in the reference."""
    bits = q_u8.to(torch.int32)
    h = ((bits & 0x80) << 8) | ((bits & 0x7F) << 7)
    return h.to(torch.int16).view(torch.float16).float() * 256.0


def dequant_w8(qweight, scales, group_n, group_k):
    """Block-scaled 8-bit weight -> fp32 (kernels/quantization/w8a16.py:100-120;
    tests/kernels/test_quantization.py:43-46,63; executor/weight_utils.py:59-72)."""
    n, k = qweight.shape
    w = fp8_e4m3_bits_to_float(qweight) if qweight.dtype == torch.uint8 else qweight.float()
    gk = min(group_k, k)
    s = scales.float()
    if s.dim() == 1:
        s = s.unsqueeze(-1)
    s = s.repeat_interleave(group_n, 0)[:n].repeat_interleave(gk, 1)[:, :k]
    return w * s


def w8a16_matmul(x, qweight, scales, group_n, group_k, bias=None):
    out = x.float() @ dequant_w8(qweight, scales, group_n, group_k).T
    if bias is not None:
        out = out + bias.float()
    return out.to(x.dtype)


def quantize_activations_int8(x):
    """Per-token dynamic quantisation exactly as the kernel does it: ``scale = absmax/127`` (1 if 0),
    ``q = trunc_toward_zero(x / scale)`` -- no rounding, no clamp (kernels/quantization/w8a8.py:50-66)."""
    xf = x.float()
    scale = xf.abs().amax(dim=-1) / 127.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    q = torch.trunc(xf / scale.unsqueeze(-1)).to(torch.int8)
    return q, scale


def smoothquant_matmul(x, qweight, weight_scales, bias=None):
    """int8 x int8 -> int32, ``acc * a_scale[m] * b_scale[n]`` (+ bias) -> x.dtype
    (kernels/quantization/w8a8.py:105-130,185-216)."""
    qa, a_scale = quantize_activations_int8(x)
    ws = weight_scales.float().reshape(-1)
    acc = qa.to(torch.int64) @ qweight.to(torch.int64).T  # exact
    out = acc.float() * a_scale.unsqueeze(-1) * ws.unsqueeze(0)
    if bias is not None:
        out = out + bias.float()
    return out.to(x.dtype)


# --------------------------------------------------------------------------- #
# MoE (kernels/fused_moe.py:45-99,352-438; tests/kernels/test_fused_moe.py:16-36)
# --------------------------------------------------------------------------- #
def moe_align_block_size(topk_ids, block_size, num_experts):
    """Sorted-slot protocol (kernels/fused_moe.py:45-99): stable sort of slot ids by expert, each
    expert run padded to ``block_size`` with the sentinel ``num_slots``; ``expert_ids`` per block;
    ``num_tokens_post_padded``. Written as explicit loops (the definition, not the torch trick)."""
    flat = topk_ids.reshape(-1).tolist()
    num_slots = len(flat)
    per_expert: list[list[int]] = [[] for _ in range(num_experts)]
    for slot, e in enumerate(flat):
        per_expert[e].append(slot)
    max_padded = num_slots + num_experts * (block_size - 1)
    sorted_ids = [num_slots] * max_padded
    max_blocks = (max_padded + block_size - 1) // block_size
    expert_ids: list[int] = []
    pos = 0
    for e in range(num_experts):
        run = per_expert[e]
        padded = (len(run) + block_size - 1) // block_size * block_size
        sorted_ids[pos : pos + len(run)] = run
        expert_ids += [e] * (padded // block_size)
        pos += padded
    num_post = pos
    # blocks past the padded total are never read; the reference fills them via searchsorted+clamp
    expert_ids += [num_experts - 1] * (max_blocks - len(expert_ids))
    return (
        torch.tensor(sorted_ids, dtype=torch.int32),
        torch.tensor(expert_ids[:max_blocks], dtype=torch.int32),
        torch.tensor([num_post], dtype=torch.int32),
    )


def fused_moe(hidden_states, w1, w2, topk_weights, topk_ids, w1_scale=None, w2_scale=None, group_n=0, group_k=0):
    """Per-expert loop with the kernel's rounding points (kernels/fused_moe.py:405-437): gate_up and
    the activation are rounded to the activation dtype, the routing weight (already cast to that
    dtype, :397) multiplies the fp32 accumulator before the cast (:201-205), and the top-k sum runs
    in fp32 over the rounded per-slot rows (:316-333)."""
    dtype = hidden_states.dtype
    num_tokens, hidden = hidden_states.shape
    num_experts, two_inter, _ = w1.shape
    inter = two_inter // 2
    top_k = topk_ids.shape[1]
    x = hidden_states.float()
    flat_w = topk_weights.reshape(-1).to(dtype).float()
    expanded = torch.zeros(num_tokens * top_k, hidden, dtype=dtype)
    for slot in range(num_tokens * top_k):
        e = int(topk_ids.reshape(-1)[slot])
        tok = slot // top_k
        if w1_scale is None:
            w1e, w2e = w1[e].float(), w2[e].float()
        else:
            w1e = dequant_w8(w1[e], w1_scale[e], group_n, group_k)
            w2e = dequant_w8(w2[e], w2_scale[e], group_n, group_k)
        gate_up = (x[tok] @ w1e.T).to(dtype).float()
        act = (F.silu(gate_up[:inter]) * gate_up[inter:]).to(dtype).float()
        expanded[slot] = ((act @ w2e.T) * flat_w[slot]).to(dtype)
    return expanded.float().view(num_tokens, top_k, hidden).sum(dim=1).to(dtype)


def route_topk(hidden_states, router_weight, top_k, norm_topk_prob=True):
    """Router (models/qwen3_moe.py:85-100): logits in the activation dtype, fp32 softmax over all
    experts, top-k, optional renormalisation, weights cast back to the activation dtype."""
    logits = F.linear(hidden_states.float(), router_weight.float()).to(hidden_states.dtype)
    probs = torch.softmax(logits.float(), dim=-1)
    w, ids = torch.topk(probs, top_k, dim=-1)
    if norm_topk_prob:
        w = w / w.sum(dim=-1, keepdim=True)
    return w.to(hidden_states.dtype), ids


# --------------------------------------------------------------------------- #
# dense projection (torch F.linear: models/quantization/methods/unquantized.py:21-22)
# --------------------------------------------------------------------------- #
def linear(x, weight, bias=None):
    out = x.float() @ weight.float().T
    if bias is not None:
        out = out + bias.float()
    return out.to(x.dtype)

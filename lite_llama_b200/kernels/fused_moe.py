"""Fused MoE: top-k routed experts as grouped tcgen05 GEMMs
(reference: lite_llama/kernels/fused_moe.py:45-99 ``moe_align_block_size``, :352-438 ``fused_moe``;
router: lite_llama/models/qwen3_moe.py:85-100)."""

from __future__ import annotations

import torch

from ._common import check, dtype_code, index_width, lib, ptr, require_cuda, stream_handle

_KIND_F16, _KIND_FP8, _KIND_INT8 = 0, 1, 2


@torch.no_grad()
def moe_align_block_size(topk_ids: torch.Tensor, block_size: int, num_experts: int):
    """Sort token slots by expert and pad each expert's run to ``block_size``.

    Returns ``(sorted_token_ids, expert_ids, num_tokens_post_padded)`` with the reference's exact
    protocol: slot ids (``token*top_k + k``) stable-sorted by expert, sentinel ``num_slots`` in the
    padding, one expert id per row block, padded total on the device. One kernel, no host sync.
    """
    require_cuda(topk_ids)
    ids = topk_ids.reshape(-1)
    if ids.dtype not in (torch.int32, torch.int64):
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    num_slots = ids.numel()
    max_padded = num_slots + num_experts * (block_size - 1)
    max_blocks = (max_padded + block_size - 1) // block_size
    dev = topk_ids.device
    sorted_ids = torch.empty(max_padded, dtype=torch.int32, device=dev)
    expert_ids = torch.empty(max_blocks, dtype=torch.int32, device=dev)
    num_post = torch.empty(1, dtype=torch.int32, device=dev)
    check(
        lib().b2l_moe_align_block_size(
            ids.data_ptr(), index_width(ids, "topk_ids"), num_slots, int(block_size), int(num_experts),
            sorted_ids.data_ptr(), expert_ids.data_ptr(), num_post.data_ptr(), stream_handle(),
        )
    )
    return sorted_ids, expert_ids, num_post


def _quant_kind(weight: torch.Tensor, scale: torch.Tensor | None) -> int:
    if scale is None:
        return _KIND_F16
    if weight.dtype == torch.uint8:
        return _KIND_FP8
    if weight.dtype == torch.int8:
        return _KIND_INT8
    raise ValueError(f"quantised expert weights must be uint8 or int8, got {weight.dtype}")


def _block_size(num_tokens: int, top_k: int = 1, num_experts: int = 1) -> int:
    """Row-block of the grouped GEMM (= UMMA N): 16 / 32 / 64. The reference picks BLOCK_M from the token count
    alone (fused_moe.py:214-222); what a block holds is one expert's slots, so the choice here follows the expected
    slots per expert (tokens * top_k / experts): 32 decode tokens over 128 experts are ~2 slots per expert, and a
    16-row block halves the padded rows and lets two CTAs share an SM (csrc/moe_gemm.cu). The result does not
    depend on the block size (every slot is the same dot product)."""
    per_expert = num_tokens * top_k / max(1, num_experts)
    if num_tokens <= 16 or per_expert <= 8:
        return 16
    if num_tokens <= 64 or per_expert <= 24:
        return 32
    return 64


def _moe_gemm(a, w, c, scale, kind, topk_weights, sorted_ids, expert_ids, num_post, top_k, block, group_n, group_k):
    num_experts, n, k = w.shape
    if w.stride(-1) != 1 or a.stride(-1) != 1:
        raise ValueError("fused_moe: last dims must be contiguous")
    if w.stride(0) != n * w.stride(1):
        w = w.contiguous()
    se = sn = sk = 0
    if scale is not None:
        if scale.dtype != torch.float32:
            scale = scale.float()
        se, sn, sk = scale.stride()
    check(
        lib().b2l_moe_gemm(
            c.data_ptr(), a.data_ptr(), w.data_ptr(), kind, ptr(scale), ptr(topk_weights), sorted_ids.data_ptr(),
            expert_ids.data_ptr(), num_post.data_ptr(), sorted_ids.numel(), c.shape[0], int(top_k), int(block),
            num_experts, n, k, a.stride(0), w.stride(1), c.stride(0), int(group_n or 1), int(group_k or k),
            se, sn, sk, stream_handle(),
        )
    )


@torch.no_grad()
def fused_moe(
    hidden_states: torch.Tensor,
    w1: torch.Tensor,
    w2: torch.Tensor,
    topk_weights: torch.Tensor,
    topk_ids: torch.Tensor,
    *,
    w1_scale: torch.Tensor | None = None,
    w2_scale: torch.Tensor | None = None,
    group_n: int = 0,
    group_k: int = 0,
) -> torch.Tensor:
    """``sum_k w_k * (silu(x @ W1g.T) * (x @ W1u.T)) @ W2.T`` over the routed experts.

    Same arguments, rounding points (gate_up, activation and the per-slot down output are rounded to
    fp16; the routing weight multiplies in fp32 before that cast; the top-k sum accumulates in fp32)
    and errors as the reference.
    """
    require_cuda(hidden_states, w1, w2, topk_weights, topk_ids, w1_scale, w2_scale)
    if hidden_states.dtype != torch.float16:
        raise ValueError(f"fused_moe activations must be fp16, got {hidden_states.dtype}")
    num_tokens, hidden = hidden_states.shape
    num_experts, two_inter, _ = w1.shape
    intermediate = two_inter // 2
    top_k = topk_ids.shape[1]
    device, dtype = hidden_states.device, hidden_states.dtype
    kind = _quant_kind(w1, w1_scale)
    if kind != _quant_kind(w2, w2_scale):
        raise ValueError("w1 and w2 must use the same quantisation format")
    # the reference asks for multiples of 128 (fused_moe.py:391-394); this kernel closes an accumulator segment every
    # group_k elements and takes any multiple of 16, which is what sub-block tensor-parallel shards need (TP = 4 of
    # Qwen3-30B-A3B: 192 columns per rank -> 64-wide scale groups)
    if kind and group_k % 16 != 0 and group_k < min(hidden, intermediate):
        raise ValueError(f"group_k ({group_k}) must be a multiple of 16 unless it covers K")
    if kind == _KIND_F16 and w1.dtype != dtype:
        raise ValueError("fp16 expert weights must match the activation dtype")
    x = hidden_states if hidden_states.stride(-1) == 1 and hidden_states.stride(0) % 8 == 0 else hidden_states.contiguous()
    flat_weights = topk_weights.reshape(-1).to(dtype).contiguous()
    block = _block_size(num_tokens, top_k, num_experts)
    sorted_ids, expert_ids, num_post = moe_align_block_size(topk_ids, block, num_experts)

    gate_up = torch.empty((num_tokens * top_k, two_inter), device=device, dtype=dtype)
    _moe_gemm(x, w1, gate_up, w1_scale, kind, None, sorted_ids, expert_ids, num_post, top_k, block, group_n, group_k)

    act = torch.empty((num_tokens * top_k, intermediate), device=device, dtype=dtype)
    check(
        lib().b2l_swiglu(
            act.data_ptr(), gate_up.data_ptr(), gate_up[:, intermediate:].data_ptr(), num_tokens * top_k, intermediate,
            gate_up.stride(0), gate_up.stride(0), act.stride(0), dtype_code(act, "act"), stream_handle(),
        )
    )

    expanded = torch.empty((num_tokens * top_k, hidden), device=device, dtype=dtype)
    _moe_gemm(act, w2, expanded, w2_scale, kind, flat_weights, sorted_ids, expert_ids, num_post, 1, block, group_n,
              group_k)

    out = torch.empty((num_tokens, hidden), device=device, dtype=dtype)
    check(lib().b2l_moe_sum(out.data_ptr(), expanded.data_ptr(), num_tokens, hidden, top_k, dtype_code(out, "out"),
                            stream_handle()))
    return out


@torch.no_grad()
def route_topk(router_logits: torch.Tensor, top_k: int, renormalize: bool = True):
    """``(weights [T, k] in the logits dtype, ids [T, k] int64)``: fp32 softmax over all experts,
    top-k, optional renormalisation (lite_llama/models/qwen3_moe.py:94-100), one kernel."""
    require_cuda(router_logits)
    if router_logits.stride(-1) != 1:
        router_logits = router_logits.contiguous()
    tokens, num_experts = router_logits.shape
    w = torch.empty((tokens, top_k), dtype=router_logits.dtype, device=router_logits.device)
    ids = torch.empty((tokens, top_k), dtype=torch.int64, device=router_logits.device)
    check(
        lib().b2l_moe_route_topk(
            w.data_ptr(), ids.data_ptr(), router_logits.data_ptr(), router_logits.stride(0), tokens, num_experts,
            int(top_k), 1 if renormalize else 0, dtype_code(router_logits, "router_logits"), stream_handle(),
        )
    )
    return w, ids

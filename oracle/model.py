"""CPU restatement of one decode step of the reference decoder, composed from ``oracle/ops.py``.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ops.py header). Used (1) by tests to check the host
layer (``lite_llama_b200.executor``) end to end on tiny geometries and (2) by ``bench.py`` as the
``cpu_baseline`` / ``--impl reference`` leg: the reference has no CPU implementation of its Triton
path, so this port of its forward (lite_llama/models/base.py:204-245,299-319,447-489) in fp32 torch
is what runs on the host cores. Quantised projections are dequantised ONCE up front (an fp32 CPU
model, like HF transformers on CPU would hold) -- that favours the CPU side.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import ops


@dataclass
class CpuLayer:
    input_norm: torch.Tensor
    post_norm: torch.Tensor
    wq: torch.Tensor      # fp32 [q_size, hidden]      (dequantised when the scheme is quantised)
    wkv: torch.Tensor     # fp32 [2*kv_size, hidden]   (k rows then v rows: the reference's fused kv_proj)
    wo: torch.Tensor
    wgate: torch.Tensor
    wup: torch.Tensor
    wdown: torch.Tensor
    bq: torch.Tensor | None = None
    bkv: torch.Tensor | None = None
    q_norm: torch.Tensor | None = None
    k_norm: torch.Tensor | None = None
    moe: dict | None = None   # routed-expert FFN instead of wgate/wup/wdown (lite_llama/models/qwen3_moe.py:48-122)


def _maybe_quant(w: torch.Tensor, method: str, group_size: int = 128) -> torch.Tensor:
    """fp16-rounded weight -> the fp32 matrix the scheme's GEMM effectively multiplies by."""
    w16 = w.to(torch.float16)
    if method == "fp16":
        return w16.float()
    if method == "int4":
        q, s, z = ops.quantize_int4_groupwise(w16.float(), group_size)
        return ops.dequant_int4(q, s, z, group_size)
    if method in ("int8", "smoothquant"):
        q, s = ops.quantize_int8_per_channel(w16.float())
        return q.float() * s
    if method == "fp8":
        q, s = ops.quantize_fp8_per_channel(w16.float())
        return ops.fp8_e4m3_bits_to_float(q) * s
    raise ValueError(method)


def random_cpu_layers(cfg, method: str, num_layers: int, seed: int = 0, draw=None):
    """Same draw order and scaling as ``lite_llama_b200.executor.decoder.random_weights`` (TP=1), so a
    CPU generator seeded identically reproduces the GPU model's weights for tiny test geometries."""
    gen = torch.Generator().manual_seed(seed)
    if draw is None:
        def draw(*shape):
            return torch.randn(*shape, generator=gen, dtype=torch.float32) * 0.02
    h = cfg.hidden_size
    embed = draw(cfg.vocab_size, h).to(torch.float16)
    lm_head = embed if cfg.tie_word_embeddings else draw(cfg.vocab_size, h).to(torch.float16)
    layers = []
    for _ in range(num_layers):
        wq, wk, wv = draw(cfg.q_size, h), draw(cfg.kv_size, h), draw(cfg.kv_size, h)
        wo = draw(h, cfg.q_size)
        is_moe = getattr(cfg, "num_experts", 0) > 0
        if not is_moe:
            wg, wu = draw(cfg.intermediate_size, h), draw(cfg.intermediate_size, h)
            wd = draw(h, cfg.intermediate_size)
        bq = bkv = None
        if cfg.qkv_bias:
            bq_, bk_, bv_ = draw(cfg.q_size), draw(cfg.kv_size), draw(cfg.kv_size)
            bq, bkv = bq_.to(torch.float16).float(), torch.cat([bk_, bv_]).to(torch.float16).float()
        # the GPU side quantises the row-concatenated [q|k|v] and [gate|up] matrices; per-row /
        # per-(row, group) schemes make that identical to quantising each piece on its own
        moe = None
        if is_moe:
            e, im = cfg.num_experts, cfg.moe_intermediate_size
            router = draw(e, h).to(torch.float16)
            gate, up, down = draw(e, im, h), draw(e, im, h), draw(e, h, im)
            moe = dict(router=router, top_k=cfg.num_experts_per_tok, norm=cfg.norm_topk_prob, group_n=0, group_k=0,
                       w1_scale=None, w2_scale=None)
            if method == "fp8":    # checkpoint format: 128 x 128 block scales over the full expert matrices
                qg, sg = ops.quantize_fp8_blockwise(gate)
                qu, su = ops.quantize_fp8_blockwise(up)
                qd, sd = ops.quantize_fp8_blockwise(down)
                moe.update(w1=torch.cat([qg, qu], dim=1), w1_scale=torch.cat([sg, su], dim=1), w2=qd, w2_scale=sd,
                           group_n=128, group_k=128)
            else:
                moe.update(w1=torch.cat([gate, up], dim=1).to(torch.float16), w2=down.to(torch.float16))
            wg = wu = wd = None
        layers.append(CpuLayer(
            input_norm=torch.ones(h, dtype=torch.float16), post_norm=torch.ones(h, dtype=torch.float16),
            wq=_maybe_quant(wq, method), wkv=_maybe_quant(torch.cat([wk, wv]), method), wo=_maybe_quant(wo, method),
            wgate=None if is_moe else _maybe_quant(wg, method), wup=None if is_moe else _maybe_quant(wu, method),
            wdown=None if is_moe else _maybe_quant(wd, method), moe=moe,
            bq=bq, bkv=bkv,
            q_norm=torch.ones(cfg.head_dim, dtype=torch.float16) if cfg.use_qk_norm else None,
            k_norm=torch.ones(cfg.head_dim, dtype=torch.float16) if cfg.use_qk_norm else None,
        ))
    return embed, lm_head, layers


def rope_tables(cfg, max_pos: int):
    dim = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    freqs = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(torch.float16), emb.sin().to(torch.float16)


def _lin(x16: torch.Tensor, w32: torch.Tensor, b32: torch.Tensor | None = None) -> torch.Tensor:
    out = x16.float() @ w32.T
    if b32 is not None:
        out = out + b32
    return out.to(torch.float16)


def smoothquant_lin(x16, w, bias=None):
    """W8A8 projection for the CPU model: dynamic per-token int8 activations (ops.quantize_activations_int8)
    against the already-dequantised int8 weight."""
    qa, s = ops.quantize_activations_int8(x16)
    out = (qa.float() * s.unsqueeze(-1)) @ w.T
    if bias is not None:
        out = out + bias
    return out.to(torch.float16)


class CpuDecoder:
    """fp16-storage / fp32-math forward with the reference's rounding points; one KV pool per layer
    in the reference layout ``[max_tokens, 2*Hkv, D]``."""

    def __init__(self, cfg, embed, lm_head, layers, max_tokens: int, max_pos: int, method: str = "fp16"):
        self.cfg, self.embed, self.lm_head, self.layers = cfg, embed, lm_head, layers
        self.kv = [torch.zeros(max_tokens, 2 * cfg.num_kv_heads, cfg.head_dim, dtype=torch.float16) for _ in layers]
        self.cos, self.sin = rope_tables(cfg, max_pos)
        self.final_norm = torch.ones(cfg.hidden_size, dtype=torch.float16)
        self.scale = 1.0 / math.sqrt(cfg.head_dim)
        self.lin = smoothquant_lin if method == "smoothquant" else _lin

    def forward(self, input_ids, position_ids, select_index, table, b_req_idx, b_seq_len, b_start_loc=None):
        """``input_ids/position_ids [B, S]``; S > 1 = prefill over the padded rectangle. Returns fp16
        logits ``[B, S, V]``."""
        cfg = self.cfg
        B, S = input_ids.shape
        T = B * S
        hidden = self.embed[input_ids]                                   # [B, S, h] fp16
        cos, sin = self.cos[position_ids], self.sin[position_ids]       # [B, S, D]
        residual = None
        for li, L in enumerate(self.layers):
            hidden, residual = ops.skip_rmsnorm(hidden, residual, L.input_norm, cfg.rms_norm_eps)
            x = hidden.reshape(T, cfg.hidden_size)
            xq = self.lin(x, L.wq, L.bq).view(T, cfg.num_heads, cfg.head_dim)
            xkv = self.lin(x, L.wkv, L.bkv)
            xk = xkv[:, : cfg.kv_size].reshape(T, cfg.num_kv_heads, cfg.head_dim)
            xv = xkv[:, cfg.kv_size :].reshape(T, cfg.num_kv_heads, cfg.head_dim)
            if L.q_norm is not None:
                xq = ops.rmsnorm(xq, L.q_norm, cfg.rms_norm_eps)
                xk = ops.rmsnorm(xk, L.k_norm, cfg.rms_norm_eps)
            xq = ops.rope_half_split(xq, cos, sin, B, S)
            xk = ops.rope_half_split(xk, cos, sin, B, S)
            ops.update_kv_buffer(torch.cat([xk, xv], dim=1), select_index, self.kv[li])
            if S > 1:
                attn = ops.varlen_causal_attention(xq, xk, xv, b_start_loc, b_seq_len, self.scale)
            else:
                attn = ops.paged_decode_attention(xq, self.kv[li][:, : cfg.num_kv_heads], self.kv[li][:, cfg.num_kv_heads :],
                                                  table, b_req_idx, b_seq_len, self.scale)
            o = self.lin(attn.reshape(T, cfg.q_size), L.wo)
            hidden, residual = ops.skip_rmsnorm(o.view(B, S, -1), residual, L.post_norm, cfg.rms_norm_eps)
            x2 = hidden.reshape(T, cfg.hidden_size)
            if L.moe is not None:
                m = L.moe
                w, ids = ops.route_topk(x2, m["router"], m["top_k"], m["norm"])
                hidden = ops.fused_moe(x2, m["w1"], m["w2"], w, ids, m["w1_scale"], m["w2_scale"], m["group_n"],
                                       m["group_k"]).view(B, S, -1)
                continue
            act = ops.swiglu(self.lin(x2, L.wgate), self.lin(x2, L.wup))
            hidden = self.lin(act, L.wdown).view(B, S, -1)
        hidden, _ = ops.skip_rmsnorm(hidden, residual, self.final_norm, cfg.rms_norm_eps)
        return _lin(hidden.reshape(T, -1), self.lm_head.float()).view(B, S, -1)

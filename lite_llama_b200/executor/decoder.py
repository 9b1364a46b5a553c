"""Decoder forward as a flat per-layer op list over the B200 kernels.

Restates the reference's ``CausalLM -> DecoderLayer -> Attention/PagedAttention + FusedMLP`` call
sequence (lite_llama/models/base.py:131-142,204-245,263-264,299-319,447-489) with the same operator
order and rounding points. Differences are launch-count only and do not change any output element:
  * q_proj and the fused kv_proj share one GEMM over the row-concatenated weight ``[q | k | v]``
    (each output element is the same dot product); gate_proj/up_proj likewise ``[gate | up]``;
  * RoPE rotates q and k where they lie in that GEMM's output (no ``.contiguous()`` copy of k,
    rope_emb.py:107) and the KV scatter reads k|v from there (no ``torch.cat``, base.py:76);
  * cos/sin come from a table built once with the reference's formula
    (models/rotary_embedding.py:33-42,123-137) instead of ~8 torch ops per step.
Tensor parallelism follows lite_llama/models/weights.py:125-134: q/kv/gate/up split on output rows,
o/down on the contracted dim, one SUM all-reduce after o_proj and after down_proj.
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import torch

from ..distributed import divide, get_tp_rank, get_tp_world_size
from ..kernels import (
    flash_attention2_no_pad,
    flash_decoding,
    fused_moe,
    linear,
    route_topk,
    rope_emb_inplace,
    rope_kv_write,
    skip_rmsnorm,
    swiglu_forward,
    update_kv_buffer,
)
from ..kernels.quantization import smoothquant_matmul_prequant
from ..kernels.skip_rmsnorm import can_fuse_int8_quant, skip_rmsnorm_quant_int8
from ..kernels.tp_fused import allreduce_skip_rmsnorm
from ..quantization import LinearWeights, apply_linear, make_linear_weights
from ..quantization.params import quantize_fp8_blockwise, reblock_scales
from .model_config import ModelConfig

_LOG2E = 1.4426950408889634
# SmoothQuant: fold the per-token int8 quantisation into the norm that feeds the GEMM (B2L_FUSE_NORM_QUANT=0: two launches)
_FUSE_NORM_QUANT = os.environ.get("B2L_FUSE_NORM_QUANT", "1") != "0"


@dataclass
class MoeWeights:
    """One routed-expert FFN (reference: Qwen3MoeSparseMoeBlock, lite_llama/models/qwen3_moe.py:48-122). Experts are
    stacked; ``w1`` = gate rows then up rows of every expert. Under TP each rank keeps ``I / tp`` intermediate
    columns of EVERY expert (models/weights.py:129-130); the router is replicated and never quantised."""

    router: torch.Tensor                  # [E, hidden], activation dtype
    w1: torch.Tensor                      # [E, 2 * I_local, hidden]  fp16 | uint8 (e4m3 bits)
    w2: torch.Tensor                      # [E, hidden, I_local]
    w1_scale: torch.Tensor | None = None  # fp32 [E, ceil(2 I_local / group_n), ceil(hidden / group_k)]
    w2_scale: torch.Tensor | None = None  # fp32 [E, ceil(hidden / group_n), ceil(I_local / group_k)]
    group_n: int = 0
    group_k: int = 0
    top_k: int = 8
    norm_topk_prob: bool = True

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.router, self.w1, self.w2, self.w1_scale, self.w2_scale)
                   if t is not None)


@dataclass
class LayerWeights:
    input_norm: torch.Tensor
    post_norm: torch.Tensor
    qkv: LinearWeights
    o: LinearWeights
    gate_up: LinearWeights | None = None
    down: LinearWeights | None = None
    q_norm: torch.Tensor | None = None
    k_norm: torch.Tensor | None = None
    moe: MoeWeights | None = None


@dataclass
class DecoderWeights:
    embed: torch.Tensor
    lm_head: torch.Tensor
    final_norm: torch.Tensor
    layers: list[LayerWeights] = field(default_factory=list)

    def nbytes_per_step(self) -> int:
        """Weight bytes one decode step must stream (embedding gather excluded)."""
        total = self.lm_head.numel() * self.lm_head.element_size()
        for lw in self.layers:
            total += lw.qkv.nbytes() + lw.o.nbytes()
            if lw.moe is None:
                total += lw.gate_up.nbytes() + lw.down.nbytes()
            else:
                total += lw.moe.router.numel() * lw.moe.router.element_size()   # experts: see moe_expert_bytes()
            total += (lw.input_norm.numel() + lw.post_norm.numel()) * 2
        return total

    def moe_expert_bytes(self) -> int:
        """Bytes of ONE expert (both matrices + scales) on this rank, summed over layers / num_layers."""
        moe = [lw.moe for lw in self.layers if lw.moe is not None]
        if not moe:
            return 0
        m = moe[0]
        e = m.w1.shape[0]
        return (m.nbytes() - m.router.numel() * m.router.element_size()) // e


def make_moe_weights(cfg: ModelConfig, router, gate, up, down, method: str, tp: int, rank: int,
                     dtype: torch.dtype) -> MoeWeights:
    """Shard + store one layer's experts. ``gate``/``up`` ``[E, I, h]``, ``down`` ``[E, h, I]`` are the FULL matrices;
    fp8 follows the checkpoint format: quantise in 128 x 128 blocks over the full tensors, then cut the shard out of
    codes and scales. A shard boundary inside a 128-block (I / tp = 192 at TP = 4) is handled by refining the scale
    grid to gcd(128, I / tp) first (``reblock_scales``), which the grouped GEMM takes as group_n / group_k."""
    inter = cfg.moe_intermediate_size
    il = divide(inter, tp, "MoE intermediate")
    lo, hi = rank * il, (rank + 1) * il
    kw = dict(top_k=cfg.num_experts_per_tok, norm_topk_prob=cfg.norm_topk_prob)
    if method != "fp8":
        w1 = torch.cat([gate[:, lo:hi], up[:, lo:hi]], dim=1).to(dtype).contiguous()
        return MoeWeights(router.to(dtype), w1, down[:, :, lo:hi].to(dtype).contiguous(), **kw)
    blk = 128
    g = math.gcd(blk, il)
    qg, sg = quantize_fp8_blockwise(gate, blk)
    qu, su = quantize_fp8_blockwise(up, blk)
    qd, sd = quantize_fp8_blockwise(down, blk)
    if g != blk:        # refine every scale grid to g x g so that one (group_n, group_k) serves w1 and w2
        sg, su, sd = (reblock_scales(reblock_scales(t, blk, g, 1), blk, g, 2) for t in (sg, su, sd))
    w1 = torch.cat([qg[:, lo:hi], qu[:, lo:hi]], dim=1).contiguous()
    w1s = torch.cat([sg[:, lo // g:hi // g], su[:, lo // g:hi // g]], dim=1).contiguous()
    w2 = qd[:, :, lo:hi].contiguous()
    w2s = sd[:, :, lo // g:hi // g].contiguous()
    return MoeWeights(router.to(dtype), w1, w2, w1s, w2s, group_n=g, group_k=g, **kw)


def random_weights(cfg: ModelConfig, method: str = "fp16", device: str = "cuda", seed: int = 0,
                   dtype: torch.dtype = torch.float16, group_size: int = 128,
                   draw_device: str | None = None) -> DecoderWeights:
    """Seeded random-init weights of ``cfg``'s shapes (``randn * 0.02``, norm weights 1),
    quantised per projection with ``method``. Every rank draws the full matrices and keeps its shard,
    so TP=N computes the same function as TP=1. ``lm_head`` stays fp16 under every scheme
    (lite_llama/models/base.py:366,425-429)."""
    tp, rank = get_tp_world_size(), get_tp_rank()
    hq, hkv = divide(cfg.num_heads, tp, "attention heads"), divide(cfg.num_kv_heads, tp, "key/value heads")
    inter = divide(cfg.intermediate_size, tp, "MLP intermediate") if cfg.num_experts == 0 else 0
    d, h = cfg.head_dim, cfg.hidden_size
    draw_device = draw_device or device   # tests draw on the CPU to share weights with the oracle
    gen = torch.Generator(device=draw_device).manual_seed(seed)

    def draw(*shape):
        t = torch.randn(*shape, generator=gen, device=draw_device, dtype=torch.float32) * 0.02
        return t.to(device)

    embed = draw(cfg.vocab_size, h).to(dtype)
    lm_head = embed if cfg.tie_word_embeddings else draw(cfg.vocab_size, h).to(dtype)
    out = DecoderWeights(embed=embed, lm_head=lm_head, final_norm=torch.ones(h, device=device, dtype=dtype))
    for _ in range(cfg.num_layers):
        wq, wk, wv = draw(cfg.q_size, h), draw(cfg.kv_size, h), draw(cfg.kv_size, h)
        wo = draw(h, cfg.q_size)
        is_moe = cfg.num_experts > 0
        if not is_moe:
            wg, wu = draw(cfg.intermediate_size, h), draw(cfg.intermediate_size, h)
            wd = draw(h, cfg.intermediate_size)
        bias = None
        if cfg.qkv_bias:
            bq, bk, bv = draw(cfg.q_size), draw(cfg.kv_size), draw(cfg.kv_size)
            bias = torch.cat([bq[rank * hq * d:(rank + 1) * hq * d], bk[rank * hkv * d:(rank + 1) * hkv * d],
                              bv[rank * hkv * d:(rank + 1) * hkv * d]]).to(dtype)
        qkv = torch.cat([wq[rank * hq * d:(rank + 1) * hq * d], wk[rank * hkv * d:(rank + 1) * hkv * d],
                         wv[rank * hkv * d:(rank + 1) * hkv * d]], dim=0)
        o = wo[:, rank * hq * d:(rank + 1) * hq * d]
        if not is_moe:
            gu = torch.cat([wg[rank * inter:(rank + 1) * inter], wu[rank * inter:(rank + 1) * inter]], dim=0)
            dn = wd[:, rank * inter:(rank + 1) * inter]

        def pack(w, b=None, fuse_swiglu=False):
            w = w.contiguous()
            return make_linear_weights(w.to(dtype) if method == "fp16" else w.to(dtype).float(), method, b, group_size,
                                       fuse_swiglu=fuse_swiglu)

        lw = LayerWeights(
            input_norm=torch.ones(h, device=device, dtype=dtype), post_norm=torch.ones(h, device=device, dtype=dtype),
            qkv=pack(qkv, bias), o=pack(o),
        )
        if is_moe:
            e, im = cfg.num_experts, cfg.moe_intermediate_size
            router = draw(e, h)
            gate, up, down = draw(e, im, h), draw(e, im, h), draw(e, h, im)
            lw.moe = make_moe_weights(cfg, router, gate, up, down, method, tp, rank, dtype)
            del gate, up, down
        else:
            lw.gate_up, lw.down = pack(gu, fuse_swiglu=True), pack(dn)
            del wg, wu, wd, gu, dn
        if cfg.use_qk_norm:
            lw.q_norm = torch.ones(d, device=device, dtype=dtype)
            lw.k_norm = torch.ones(d, device=device, dtype=dtype)
        out.layers.append(lw)
        del wq, wk, wv, wo, qkv, o
    return out


def build_rope_table(cfg: ModelConfig, max_pos: int, device: str, dtype: torch.dtype):
    """``cos/sin [max_pos, head_dim]`` with the reference's default-RoPE formula
    (models/rotary_embedding.py:33-42: ``inv_freq = 1 / theta**(2i/d)``; :129-137: fp32 outer product,
    duplicated halves, cast to the activation dtype)."""
    dim = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
    pos = torch.arange(max_pos, dtype=torch.float32, device=device)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


class Decoder:
    """Forward skeleton (reference: CausalLM.forward, lite_llama/models/base.py:447-489)."""

    def __init__(self, cfg: ModelConfig, weights: DecoderWeights, max_pos: int, device: str = "cuda",
                 dtype: torch.dtype = torch.float16) -> None:
        cfg.validate()
        self.cfg = cfg
        self.w = weights
        tp = get_tp_world_size()
        self.num_heads = divide(cfg.num_heads, tp, "attention heads")
        self.num_kv_heads = divide(cfg.num_kv_heads, tp, "key/value heads")
        self.head_dim = cfg.head_dim
        self.q_size = self.num_heads * self.head_dim
        self.kv_size = self.num_kv_heads * self.head_dim
        self.inter = divide(cfg.intermediate_size, tp, "MLP intermediate") if cfg.num_experts == 0 else 0
        self.scale = 1.0 / math.sqrt(cfg.head_dim)
        self.cos_table, self.sin_table = build_rope_table(cfg, max_pos, device, dtype)

    # -- one decoder layer (base.py:299-319) ---------------------------------------------------
    def _layer(self, li: int, lw: LayerWeights, hidden, residual, atten_info, cos, sin, batch, seq_len):
        """``hidden`` of every layer but the first is the previous layer's row-parallel down_proj output, i.e. a
        per-rank PARTIAL under tensor parallelism: its all-reduce (linear.py:160-161) is fused into this layer's
        input norm, the o_proj one into the post-attention norm (kernels/tp_fused.py)."""
        eps = self.cfg.rms_norm_eps
        tokens = batch * seq_len
        # input norm (base.py:307) -> fused q|k|v projection (base.py:213-215)
        qkv, residual = self._norm_linear(hidden, residual, lw.input_norm, lw.qkv, eps, tokens)
        xq = qkv[:, : self.q_size].view(tokens, self.num_heads, self.head_dim)
        xk = qkv[:, self.q_size : self.q_size + self.kv_size].view(tokens, self.num_kv_heads, self.head_dim)
        xv = qkv[:, self.q_size + self.kv_size :].view(tokens, self.num_kv_heads, self.head_dim)
        if lw.q_norm is not None:                                                       # base.py:222-225
            xq, _ = skip_rmsnorm(xq, None, lw.q_norm, eps)
            xk, _ = skip_rmsnorm(xk, None, lw.k_norm, eps)
        kv_buffer = atten_info.kv_buffer[li]
        if self.head_dim % 4 == 0:
            # rope (base.py:228) + cat + cache write (base.py:76-79) in one launch; q/k rotate where they lie
            rope_kv_write(xq, xk, xv, cos, sin, atten_info.cur_select_index, kv_buffer, batch, seq_len)
        else:
            rope_emb_inplace(xq, xk, cos, sin, batch, seq_len)
            from ..kernels import update_kv_buffer_split

            update_kv_buffer_split(xk, xv, atten_info.cur_select_index, kv_buffer)
        if seq_len > 1:                                                                 # base.py:95-103
            attn = flash_attention2_no_pad(xq, xk, xv, self.scale * _LOG2E, atten_info.b_start_loc,
                                           atten_info.b_seq_len, atten_info.max_actual_seq_len)
        else:                                                                           # base.py:120-129
            attn = flash_decoding(xq, kv_buffer[:, : self.num_kv_heads], kv_buffer[:, self.num_kv_heads :],
                                  self.scale, atten_info.b_req_tokens_table, atten_info.b_req_idx,
                                  atten_info.b_seq_len, atten_info.max_actual_seq_len)
        o = apply_linear(attn.view(tokens, self.q_size), lw.o)                          # base.py:245 (partial)
        # all-reduce (linear.py:160-161) + residual add + post-attention norm (base.py:312) in one kernel
        if lw.moe is not None:
            hidden, residual = allreduce_skip_rmsnorm(o.view(batch, seq_len, -1), residual, lw.post_norm, eps)
            down = self._moe(hidden.view(tokens, -1), lw.moe)                           # qwen3_moe.py:102-111 (partial)
            return down.view(batch, seq_len, -1), residual
        gu, residual = self._norm_linear(o.view(batch, seq_len, -1), residual, lw.post_norm, lw.gate_up, eps, tokens)  # base.py:263-264
        # W4A16: the activation already ran in the GEMM epilogue (row-interleaved gate|up), gu is [tokens, inter]
        act = gu if lw.gate_up.swiglu else swiglu_forward(gu[:, : self.inter], gu[:, self.inter :])
        down = apply_linear(act, lw.down)                                               # partial: reduced by the next norm
        return down.view(batch, seq_len, -1), residual

    @staticmethod
    def _norm_linear(hidden, residual, norm_weight, w: LinearWeights, eps: float, tokens: int):
        """``skip_rmsnorm`` (with the pending tensor-parallel all-reduce of ``hidden`` when there is one) followed by the
        projection that consumes it; returns ``(projection [tokens, out], residual)``. SmoothQuant (TP = 1): the norm
        kernel also does the per-token int8 quantisation the GEMM needs (reference: two more launches and an fp16 round
        trip per norm, quantization/w8a8.py:35-68) -- the quantised activations are bit-identical to the unfused path."""
        n = hidden.shape[-1]
        if (w.method == "smoothquant" and _FUSE_NORM_QUANT and get_tp_world_size() <= 1 and can_fuse_int8_quant(n)
                and hidden.dtype == torch.float16):
            q, s, residual = skip_rmsnorm_quant_int8(hidden, residual, norm_weight, eps)
            return smoothquant_matmul_prequant(q, s, w.qweight, w.scales, bias=w.bias, out_dtype=hidden.dtype), residual
        if residual is None:
            hidden, residual = skip_rmsnorm(hidden, residual, norm_weight, eps)
        else:
            hidden, residual = allreduce_skip_rmsnorm(hidden, residual, norm_weight, eps)
        return apply_linear(hidden.view(tokens, n), w), residual

    @staticmethod
    def _moe(x: torch.Tensor, m: MoeWeights) -> torch.Tensor:
        """Router (fp16 logits -> fp32 softmax -> top-k -> renormalise, qwen3_moe.py:85-100) + routed SwiGLU experts
        (fused_moe). Under TP the result is this rank's partial sum over its slice of every expert's intermediate
        dimension; its all-reduce (qwen3_moe.py:110) is fused into the next norm like the dense MLP's."""
        logits = linear(x, m.router)
        weights, ids = route_topk(logits, m.top_k, m.norm_topk_prob)
        return fused_moe(x, m.w1, m.w2, weights, ids, w1_scale=m.w1_scale, w2_scale=m.w2_scale,
                         group_n=m.group_n, group_k=m.group_k)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, atten_info,
                last_token_index: torch.Tensor | None = None) -> torch.Tensor:
        """``[batch, seq_len]`` ids -> ``[batch, seq_len, vocab]`` logits (all positions, like the
        reference); with ``last_token_index [batch]`` only those positions are projected."""
        batch, seq_len = input_ids.shape
        hidden = torch.nn.functional.embedding(input_ids, self.w.embed)                 # base.py:470-472
        cos = self.cos_table[position_ids]                                              # base.py:473
        sin = self.sin_table[position_ids]
        residual = None
        for li, lw in enumerate(self.w.layers):                                         # base.py:476
            hidden, residual = self._layer(li, lw, hidden, residual, atten_info, cos, sin, batch, seq_len)
        if last_token_index is not None:
            rows = torch.arange(batch, device=hidden.device)
            hidden = hidden[rows, last_token_index].unsqueeze(1)
            residual = residual[rows, last_token_index].unsqueeze(1)
        hidden, _ = allreduce_skip_rmsnorm(hidden.contiguous(), residual.contiguous(), self.w.final_norm,
                                           self.cfg.rms_norm_eps)                       # base.py:486
        return linear(hidden, self.w.lm_head)                                           # base.py:489

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
from dataclasses import dataclass, field

import torch

from ..distributed import divide, get_tp_rank, get_tp_world_size
from ..kernels import (
    flash_attention2_no_pad,
    flash_decoding,
    linear,
    rope_emb_inplace,
    skip_rmsnorm,
    swiglu_forward,
    update_kv_buffer,
)
from ..kernels.tp_fused import allreduce_skip_rmsnorm
from ..quantization import LinearWeights, apply_linear, make_linear_weights
from .model_config import ModelConfig

_LOG2E = 1.4426950408889634


@dataclass
class LayerWeights:
    input_norm: torch.Tensor
    post_norm: torch.Tensor
    qkv: LinearWeights
    o: LinearWeights
    gate_up: LinearWeights
    down: LinearWeights
    q_norm: torch.Tensor | None = None
    k_norm: torch.Tensor | None = None


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
            total += lw.qkv.nbytes() + lw.o.nbytes() + lw.gate_up.nbytes() + lw.down.nbytes()
            total += (lw.input_norm.numel() + lw.post_norm.numel()) * 2
        return total


def random_weights(cfg: ModelConfig, method: str = "fp16", device: str = "cuda", seed: int = 0,
                   dtype: torch.dtype = torch.float16, group_size: int = 128,
                   draw_device: str | None = None) -> DecoderWeights:
    """Seeded random-init weights of ``cfg``'s shapes (``randn * 0.02``, norm weights 1),
    quantised per projection with ``method``. Every rank draws the full matrices and keeps its shard,
    so TP=N computes the same function as TP=1. ``lm_head`` stays fp16 under every scheme
    (lite_llama/models/base.py:366,425-429)."""
    tp, rank = get_tp_world_size(), get_tp_rank()
    hq, hkv = divide(cfg.num_heads, tp, "attention heads"), divide(cfg.num_kv_heads, tp, "key/value heads")
    inter = divide(cfg.intermediate_size, tp, "MLP intermediate")
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
        gu = torch.cat([wg[rank * inter:(rank + 1) * inter], wu[rank * inter:(rank + 1) * inter]], dim=0)
        dn = wd[:, rank * inter:(rank + 1) * inter]

        def pack(w, b=None):
            w = w.contiguous()
            return make_linear_weights(w.to(dtype) if method == "fp16" else w.to(dtype).float(), method, b, group_size)

        lw = LayerWeights(
            input_norm=torch.ones(h, device=device, dtype=dtype), post_norm=torch.ones(h, device=device, dtype=dtype),
            qkv=pack(qkv, bias), o=pack(o), gate_up=pack(gu), down=pack(dn),
        )
        if cfg.use_qk_norm:
            lw.q_norm = torch.ones(d, device=device, dtype=dtype)
            lw.k_norm = torch.ones(d, device=device, dtype=dtype)
        out.layers.append(lw)
        del wq, wk, wv, wo, wg, wu, wd, qkv, o, gu, dn
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
        self.inter = divide(cfg.intermediate_size, tp, "MLP intermediate")
        self.scale = 1.0 / math.sqrt(cfg.head_dim)
        self.cos_table, self.sin_table = build_rope_table(cfg, max_pos, device, dtype)

    # -- one decoder layer (base.py:299-319) ---------------------------------------------------
    def _layer(self, li: int, lw: LayerWeights, hidden, residual, atten_info, cos, sin, batch, seq_len):
        """``hidden`` of every layer but the first is the previous layer's row-parallel down_proj output, i.e. a
        per-rank PARTIAL under tensor parallelism: its all-reduce (linear.py:160-161) is fused into this layer's
        input norm, the o_proj one into the post-attention norm (kernels/tp_fused.py)."""
        eps = self.cfg.rms_norm_eps
        if residual is None:
            hidden, residual = skip_rmsnorm(hidden, residual, lw.input_norm, eps)      # base.py:307
        else:
            hidden, residual = allreduce_skip_rmsnorm(hidden, residual, lw.input_norm, eps)
        tokens = batch * seq_len
        x = hidden.view(tokens, self.cfg.hidden_size)
        qkv = apply_linear(x, lw.qkv)                                                   # base.py:213-215
        xq = qkv[:, : self.q_size].view(tokens, self.num_heads, self.head_dim)
        xk = qkv[:, self.q_size : self.q_size + self.kv_size].view(tokens, self.num_kv_heads, self.head_dim)
        xv = qkv[:, self.q_size + self.kv_size :].view(tokens, self.num_kv_heads, self.head_dim)
        if lw.q_norm is not None:                                                       # base.py:222-225
            xq, _ = skip_rmsnorm(xq, None, lw.q_norm, eps)
            xk, _ = skip_rmsnorm(xk, None, lw.k_norm, eps)
        rope_emb_inplace(xq, xk, cos, sin, batch, seq_len)                              # base.py:228
        kv_buffer = atten_info.kv_buffer[li]
        if lw.q_norm is None:
            # k|v are adjacent column blocks of qkv: one strided [tokens, 2*Hkv, D] source, no cat
            xkv = qkv[:, self.q_size :].view(tokens, 2 * self.num_kv_heads, self.head_dim)
            update_kv_buffer(xkv, atten_info.cur_select_index, kv_buffer)              # base.py:76-79
        else:
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
        hidden, residual = allreduce_skip_rmsnorm(o.view(batch, seq_len, -1), residual, lw.post_norm, eps)
        gu = apply_linear(hidden.view(tokens, -1), lw.gate_up)                          # base.py:263-264
        act = swiglu_forward(gu[:, : self.inter], gu[:, self.inter :])
        down = apply_linear(act, lw.down)                                               # partial: reduced by the next norm
        return down.view(batch, seq_len, -1), residual

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

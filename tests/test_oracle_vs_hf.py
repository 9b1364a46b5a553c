"""Pins the whole-model CPU oracle (``oracle/model.py``) against an independent implementation: HF transformers'
Llama / Qwen2 / Qwen3 decoders built from the same random weights (no checkpoint exists offline).

The reference validates itself the same way -- logits parity against HF (tests/models/test_qwen3_moe.py:185-245) and
greedy text against a recorded HF-free golden (tests/golden/cases.py:20-37).  Here: prefill logits of the oracle
(fp16 storage, fp32 math, the reference's rounding points) against HF in fp32, and token-for-token greedy decode.
"""

from __future__ import annotations

import pytest
import torch

from lite_llama_b200.executor.model_config import PRESETS
from oracle import model as om

transformers = pytest.importorskip("transformers")


def _hf_model(cfg, embed, lm_head, layers):
    common = dict(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                  num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                  num_key_value_heads=cfg.num_kv_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                  max_position_embeddings=512, tie_word_embeddings=False)
    if cfg.use_qk_norm:
        hc = transformers.Qwen3Config(head_dim=cfg.head_dim, attention_bias=False, **common)
        m = transformers.Qwen3ForCausalLM(hc)
    elif cfg.qkv_bias:
        hc = transformers.Qwen2Config(**common)
        m = transformers.Qwen2ForCausalLM(hc)
    else:
        hc = transformers.LlamaConfig(head_dim=cfg.head_dim, attention_bias=False, **common)
        m = transformers.LlamaForCausalLM(hc)
    sd = {"model.embed_tokens.weight": embed.float(), "lm_head.weight": lm_head.float(),
          "model.norm.weight": torch.ones(cfg.hidden_size)}
    for i, L in enumerate(layers):
        pre = f"model.layers.{i}."
        sd[pre + "self_attn.q_proj.weight"] = L.wq
        sd[pre + "self_attn.k_proj.weight"] = L.wkv[: cfg.kv_size]
        sd[pre + "self_attn.v_proj.weight"] = L.wkv[cfg.kv_size:]
        sd[pre + "self_attn.o_proj.weight"] = L.wo
        sd[pre + "mlp.gate_proj.weight"] = L.wgate
        sd[pre + "mlp.up_proj.weight"] = L.wup
        sd[pre + "mlp.down_proj.weight"] = L.wdown
        sd[pre + "input_layernorm.weight"] = L.input_norm.float()
        sd[pre + "post_attention_layernorm.weight"] = L.post_norm.float()
        if L.bq is not None:
            sd[pre + "self_attn.q_proj.bias"] = L.bq
            sd[pre + "self_attn.k_proj.bias"] = L.bkv[: cfg.kv_size]
            sd[pre + "self_attn.v_proj.bias"] = L.bkv[cfg.kv_size:]
        if L.q_norm is not None:
            sd[pre + "self_attn.q_norm.weight"] = L.q_norm.float()
            sd[pre + "self_attn.k_norm.weight"] = L.k_norm.float()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m.float().eval()


@pytest.mark.parametrize("name", ["tiny-llama", "tiny-qwen2", "tiny-qwen3"])
def test_cpu_oracle_matches_hf_transformers(name):
    torch.manual_seed(0)
    cfg = PRESETS[name]
    embed, lm_head, layers = om.random_cpu_layers(cfg, "fp16", cfg.num_layers, seed=3)
    hf = _hf_model(cfg, embed, lm_head, layers)
    B, S, NEW = 2, 12, 6
    max_seq = S + NEW
    dec = om.CpuDecoder(cfg, embed, lm_head, layers, max_tokens=B * max_seq, max_pos=max_seq)
    ids = torch.randint(0, cfg.vocab_size, (B, S))
    table = torch.arange(B * max_seq, dtype=torch.int32).view(B, max_seq)
    req = torch.arange(B, dtype=torch.int32)
    # ---- prefill: all positions' logits ----
    pos = torch.arange(S).expand(B, S)
    logits = dec.forward(ids, pos, table[:, :S].reshape(-1), table, req, torch.full((B,), S, dtype=torch.int32),
                         b_start_loc=torch.arange(B, dtype=torch.int32) * S)
    with torch.no_grad():
        ref = hf(input_ids=ids).logits
    # the oracle rounds activations to fp16 between ops like the reference does; HF runs fp32 end to end
    torch.testing.assert_close(logits.float(), ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    assert (logits[:, -1].float().argmax(-1) == ref[:, -1].argmax(-1)).all()
    # ---- greedy decode, token for token ----
    with torch.no_grad():
        hf_tokens = hf.generate(ids, max_new_tokens=NEW, do_sample=False, use_cache=True)[:, S:]
    nxt = logits[:, -1].float().argmax(-1)
    mine = [nxt]
    for step in range(1, NEW):
        cur = S + step - 1
        lg = dec.forward(nxt.view(B, 1), torch.full((B, 1), cur), table[:, cur].contiguous(), table, req,
                         torch.full((B,), cur + 1, dtype=torch.int32))
        nxt = lg[:, 0].float().argmax(-1)
        mine.append(nxt)
    assert torch.equal(torch.stack(mine, dim=1), hf_tokens)

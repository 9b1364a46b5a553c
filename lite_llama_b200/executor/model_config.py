"""Model geometry for the decoder families the reference serves (LLaMA / Qwen2 / Qwen3 dense).

The reference wraps HF ``AutoConfig`` (lite_llama/models/config.py:48-218); no checkpoints or network
exist in this environment, so the public HF geometries are written down here.
"""

from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass(frozen=True)
class ModelConfig:
    name: str
    hidden_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    intermediate_size: int
    vocab_size: int
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-5
    qkv_bias: bool = False
    use_qk_norm: bool = False
    tie_word_embeddings: bool = False
    max_seq_len: int = 2048
    # routed-expert FFN (reference: lite_llama/models/qwen3_moe.py:48-122); 0 experts = dense SwiGLU MLP
    num_experts: int = 0
    num_experts_per_tok: int = 0
    moe_intermediate_size: int = 0
    norm_topk_prob: bool = True

    @property
    def q_size(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_kv_heads * self.head_dim

    def validate(self) -> None:
        if self.head_dim % 8 != 0:
            raise ValueError(f"head_dim {self.head_dim} must be a multiple of 8")
        if self.num_heads % self.num_kv_heads != 0:
            raise ValueError("num_heads must be a multiple of num_kv_heads")

    def with_(self, **kw) -> "ModelConfig":
        return replace(self, **kw)


PRESETS: dict[str, ModelConfig] = {
    "llama-3-8b": ModelConfig("llama-3-8b", 4096, 32, 32, 8, 128, 14336, 128256, rope_theta=5e5, rms_norm_eps=1e-5),
    "qwen2.5-7b": ModelConfig("qwen2.5-7b", 3584, 28, 28, 4, 128, 18944, 152064, rope_theta=1e6, rms_norm_eps=1e-6,
                              qkv_bias=True),
    "qwen2.5-0.5b": ModelConfig("qwen2.5-0.5b", 896, 24, 14, 2, 64, 4864, 151936, rope_theta=1e6, rms_norm_eps=1e-6,
                                qkv_bias=True, tie_word_embeddings=True),
    "qwen3-8b": ModelConfig("qwen3-8b", 4096, 36, 32, 8, 128, 12288, 151936, rope_theta=1e6, rms_norm_eps=1e-6,
                            use_qk_norm=True),
    "qwen3-30b-a3b": ModelConfig("qwen3-30b-a3b", 2048, 48, 32, 4, 128, 6144, 151936, rope_theta=1e6, rms_norm_eps=1e-6,
                                 use_qk_norm=True, num_experts=128, num_experts_per_tok=8, moe_intermediate_size=768),
    # small geometries for tests (same code paths, seconds on CPU for the oracle)
    "tiny-llama": ModelConfig("tiny-llama", 256, 2, 8, 2, 32, 512, 1000, rope_theta=1e4),
    "tiny-qwen3": ModelConfig("tiny-qwen3", 256, 2, 8, 4, 32, 512, 1000, rope_theta=1e4, qkv_bias=False,
                              use_qk_norm=True, rms_norm_eps=1e-6),
    "tiny-qwen3-moe": ModelConfig("tiny-qwen3-moe", 256, 2, 8, 4, 32, 512, 1000, rope_theta=1e4, use_qk_norm=True,
                                  rms_norm_eps=1e-6, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128),
    "tiny-qwen2": ModelConfig("tiny-qwen2", 256, 2, 8, 2, 32, 512, 1000, rope_theta=1e4, qkv_bias=True,
                              rms_norm_eps=1e-6),
}

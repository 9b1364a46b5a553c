"""Host-side mirror of the reference's executor surface (lite_llama/executor/*): the per-step
attention metadata, the KV pool + bump allocator, and the ModelRunner that drives the B200 kernels."""

from .attention_metadata import AttentionMetadata
from .kv_cache_manager import KVCacheManager
from .model_config import ModelConfig, PRESETS
from .model_runner import ModelRunner

__all__ = ["AttentionMetadata", "KVCacheManager", "ModelConfig", "PRESETS", "ModelRunner"]

"""Bind the B200 kernels into the UNMODIFIED reference package.

The reference has no plugin mechanism: its operator API is Python name binding at import time
(lite_llama/models/base.py:31-38, lite_llama/executor/model_runner.py:22,
lite_llama/models/quantization/methods/{w4a16,w8a16,w8a8,unquantized}.py). ``bind()`` rebinds those names to this
package's ctypes shims -- same signatures, argument meaning, in-place behaviour and error types -- so that the
reference's own ``LLMEngine`` / ``TextGenerator`` / ``ContinuousBatchingEngine`` and model classes run on
``libb200lite.so`` without a source change; ``unbind()`` restores the Triton kernels (A/B runs in one process).
"""

from __future__ import annotations

import importlib

from . import kernels as _k

# (module, attribute) -> replacement
_BINDINGS = [
    ("lite_llama.models.base", "flash_attention2_no_pad"),
    ("lite_llama.models.base", "flash_decoding"),
    ("lite_llama.models.base", "rope_emb_forward"),
    ("lite_llama.models.base", "skip_rmsnorm"),
    ("lite_llama.models.base", "swiglu_forward"),
    ("lite_llama.models.base", "update_kv_buffer"),
    ("lite_llama.executor.model_runner", "update_kv_index"),
    ("lite_llama.models.quantization.methods.w4a16", "w4a16_matmul"),
    ("lite_llama.models.quantization.methods.w8a16", "w8a16_matmul"),
    ("lite_llama.models.quantization.methods.w8a16", "fused_moe"),
    ("lite_llama.models.quantization.methods.w8a8", "smoothquant_matmul"),
    ("lite_llama.models.quantization.methods.unquantized", "fused_moe"),
]
_saved: dict[tuple[str, str], object] = {}


def bind(dense_linear: bool = False) -> list[str]:
    """Point the reference's kernel names at the B200 shims; returns the ``module.name`` list that was bound.
    ``dense_linear=True`` also routes ``UnquantizedLinearMethod.apply`` (``F.linear``, cuBLAS in the reference:
    methods/unquantized.py:21-22) through the tcgen05 dense GEMM."""
    done = []
    for mod_name, attr in _BINDINGS:
        mod = importlib.import_module(mod_name)
        if not hasattr(mod, attr):
            continue
        _saved.setdefault((mod_name, attr), getattr(mod, attr))
        setattr(mod, attr, getattr(_k, attr))
        done.append(f"{mod_name}.{attr}")
    if dense_linear:
        unq = importlib.import_module("lite_llama.models.quantization.methods.unquantized")
        cls = unq.UnquantizedLinearMethod
        _saved.setdefault(("unquantized", "apply"), cls.apply)

        def apply(self, layer, x):  # same contract as the reference method
            return _k.linear(x, layer.weight, layer.bias)

        cls.apply = apply
        done.append("UnquantizedLinearMethod.apply")
    return done


def unbind() -> None:
    for (mod_name, attr), fn in list(_saved.items()):
        if mod_name == "unquantized":
            unq = importlib.import_module("lite_llama.models.quantization.methods.unquantized")
            unq.UnquantizedLinearMethod.apply = fn
        else:
            setattr(importlib.import_module(mod_name), attr, fn)
    _saved.clear()

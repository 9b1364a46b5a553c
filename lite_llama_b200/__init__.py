"""lite_llama_b200 -- a B200 (sm_100a) native kernel layer behind lite_llama's op signatures.

``lite_llama_b200.kernels`` exports the same callables as ``lite_llama.kernels``
(reference: lite_llama/kernels/__init__.py:11-39); every one of them is a thin ctypes shim over
the hand-written CUDA in ``csrc/`` (built in-tree into ``libb200lite.so``). There is no Triton,
no CPU fallback and no alternative backend: if the shared library is missing, importing the
kernels raises.
"""

__version__ = "0.1.0"

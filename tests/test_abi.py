"""C-ABI surface checks that need no GPU: the in-tree library loads, exports every symbol declared
in include/b200_lite_kernels.h, and validates arguments before touching the device."""

from __future__ import annotations

import ctypes
import re
import subprocess

import pytest

from lite_llama_b200 import _lib


def test_header_parses_and_declares_the_hot_path_ops():
    protos = _lib.parse_header()
    for name in (
        "b2l_skip_rmsnorm", "b2l_rope_emb", "b2l_swiglu", "b2l_update_kv_buffer", "b2l_update_kv_index",
        "b2l_flash_decoding", "b2l_flash_attention2_no_pad", "b2l_linear", "b2l_w4a16_matmul",
        "b2l_w8a16_matmul", "b2l_quantize_activations_int8", "b2l_w8a8_gemm", "b2l_moe_sum",
    ):
        assert name in protos, f"{name} missing from the header"
    # every entry point that launches work takes the stream last
    for name, (ret, args) in protos.items():
        if ret == "int" and name not in ("b2l_version", "b2l_device_info", "b2l_set_pdl", "b2l_hang_report", "b2l_tp_comm_handle_bytes",
                                          "b2l_tp_comm_get_handle", "b2l_tp_comm_open", "b2l_tp_comm_destroy"):
            assert args[-1] == "void*", f"{name} must end with the stream argument"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _lib.parse_header():
        assert hasattr(lib, name), f"libb200lite.so does not export {name}"
    assert lib.b2l_version() >= 100


def test_library_is_sm100a_native_code():
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\w+", out.stdout))
    assert archs == {"sm_100a"}, f"unexpected code objects: {archs}"


def test_argument_validation_happens_on_the_host():
    lib = _lib.load()
    # null pointers are rejected before any CUDA call (works without a GPU)
    rc = lib.b2l_skip_rmsnorm(None, None, None, None, 1, 8, 8, 8, 8, ctypes.c_float(1e-5), 0, None)
    assert rc == -1 and b"null pointer" in lib.b2l_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.b2l_update_kv_index(p, p, 3, p, 4, p, 4, 1, 1, 1, None)
    assert rc == -1 and b"widths" in lib.b2l_last_error()
    rc = lib.b2l_flash_decoding(p, p, p, p, p, p, 4, p, 4, 1, 6, 4, 64, 16, ctypes.c_float(0.1),
                                64, 64, 64, 64, 64, 64, 1, 1, 64, 64, None, 0, None, 0, 0, None)
    assert rc == -1 and b"multiple of num_kv_heads" in lib.b2l_last_error()
    rc = lib.b2l_w4a16_matmul(p, p, p, p, p, None, 0, 1, 8, 100, 128, 8, 8, 1, 1, 8, None, 0, None, 0, None)
    assert rc == -1 and b"w4a16_matmul" in lib.b2l_last_error()


def test_workspace_queries_are_pure_host_functions():
    lib = _lib.load()
    assert lib.b2l_flash_decoding_workspace_bytes(0, 32, 8, 128, 256) == 0
    assert lib.b2l_linear_num_counters(32, 4096) == 32
    assert lib.b2l_quant_gemm_num_counters(32, 4096) == 32


def test_check_raises_value_error_like_the_reference():
    with pytest.raises(ValueError, match="b200lite error"):
        _lib.check(-1)


def test_shims_refuse_cpu_tensors():
    import torch

    from lite_llama_b200.kernels import skip_rmsnorm, w4a16_matmul

    x = torch.zeros(2, 64, dtype=torch.float16)
    with pytest.raises(ValueError, match="no CPU fallback"):
        skip_rmsnorm(x, None, torch.ones(64, dtype=torch.float16))
    # dtype errors mirror the reference's messages (kernels/quantization/w4a16.py:174-177)
    with pytest.raises(ValueError, match="activations must be fp16"):
        w4a16_matmul(x.float(), torch.zeros(8, 8, dtype=torch.int32), torch.ones(8, 1), torch.zeros(8, 1))
    with pytest.raises(ValueError, match="qweight must be int32"):
        w4a16_matmul(x, torch.zeros(8, 8, dtype=torch.int8), torch.ones(8, 1), torch.zeros(8, 1))

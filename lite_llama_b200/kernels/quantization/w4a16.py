"""W4A16 GEMM (reference: lite_llama/kernels/quantization/w4a16.py:152-207)."""

from __future__ import annotations

import torch

from .._common import check, counters, lib, ptr, require_cuda, stream_handle, zeroed_workspace

_Q_INT4 = 0


@torch.no_grad()
def w4a16_matmul(
    x: torch.Tensor,
    qweight: torch.Tensor,
    scales: torch.Tensor,
    zeros: torch.Tensor,
    *,
    group_size: int = 128,
    bias: torch.Tensor | None = None,
) -> torch.Tensor:
    """``x @ dequant(qweight).T (+ bias)``; ``qweight [N, K/8]`` int32 (nibble j of word i = K index
    8i+j), ``scales``/``zeros`` fp32 ``[N, K/group_size]``. Raises like the reference on bad dtypes."""
    if x.dtype != torch.float16:
        raise ValueError(f"w4a16 activations must be fp16, got {x.dtype}")
    if qweight.dtype != torch.int32:
        raise ValueError(f"qweight must be int32 (packed int4), got {qweight.dtype}")
    require_cuda(x, qweight, scales, zeros, bias)
    n, k_packed = qweight.shape
    k = k_packed * 8
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    if k % group_size != 0:
        raise ValueError(f"K ({k}) must be a multiple of group_size ({group_size})")
    if scales.dtype != torch.float32 or zeros.dtype != torch.float32:
        scales, zeros = scales.float(), zeros.float()
    if zeros.stride() != scales.stride():
        zeros = zeros.contiguous()
        scales = scales.contiguous()
    if qweight.stride(-1) != 1:
        qweight = qweight.contiguous()
    leading = x.shape[:-1]
    a = x.reshape(-1, k)
    if a.stride(-1) != 1 or a.stride(0) % 8 != 0 or a.data_ptr() % 16 != 0:
        a = a.contiguous()
    m = a.shape[0]
    out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    bias_f32 = 0
    if bias is not None:
        if bias.dtype == torch.float32:
            bias_f32 = 1
        elif bias.dtype != torch.float16:
            bias = bias.float()
            bias_f32 = 1
        bias = bias.contiguous()
    L = lib()
    nbytes = L.b2l_quant_gemm_workspace_bytes(_Q_INT4, m, n, k, int(group_size))
    ws = zeroed_workspace(nbytes, x.device)
    cnt = counters(x.device)
    check(
        L.b2l_w4a16_matmul(
            out.data_ptr(), a.data_ptr(), qweight.data_ptr(), scales.data_ptr(), zeros.data_ptr(), ptr(bias),
            bias_f32, m, n, k, int(group_size), a.stride(0), qweight.stride(0), scales.stride(0), scales.stride(1),
            out.stride(0), ptr(ws), nbytes, cnt.data_ptr(), cnt.numel(), stream_handle(),
        )
    )
    return out.reshape(*leading, n)


class PackedW4:
    """The three W4A16 tensors after the load-time repack into the kernel-native layout
    (``include/b200_lite_kernels.h``: b2l_w4a16_pack_*). Bit-exact permutations of the reference's tensors
    (lite_llama/kernels/quantization/w4a16.py:9-12); produced once per projection by :func:`w4a16_pack`."""

    __slots__ = ("qweight", "scales", "zeros", "n", "k", "group_size", "zeros_foldable")

    def __init__(self, qweight, scales, zeros, n, k, group_size, zeros_foldable):
        self.qweight, self.scales, self.zeros = qweight, scales, zeros
        self.n, self.k, self.group_size, self.zeros_foldable = n, k, group_size, zeros_foldable

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.qweight, self.scales, self.zeros))


@torch.no_grad()
def w4a16_pack(qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, *, group_size: int = 128) -> PackedW4:
    """Repack reference-layout ``(qweight [N, K/8] int32, scales/zeros [N, K/g] fp32)`` for
    :func:`w4a16_matmul_packed`. Load-time only (it synchronises once to classify the zero points)."""
    if qweight.dtype != torch.int32:
        raise ValueError(f"qweight must be int32 (packed int4), got {qweight.dtype}")
    require_cuda(qweight, scales, zeros)
    n, k8 = qweight.shape
    k = k8 * 8
    if k % group_size != 0:
        raise ValueError(f"K ({k}) must be a multiple of group_size ({group_size})")
    groups = k // group_size
    if tuple(scales.shape) != (n, groups) or tuple(zeros.shape) != (n, groups):
        raise ValueError(f"scales/zeros must be [{n}, {groups}], got {tuple(scales.shape)} / {tuple(zeros.shape)}")
    scales, zeros = scales.float(), zeros.float()
    if qweight.stride(-1) != 1:
        qweight = qweight.contiguous()
    L = lib()
    tiles = (n + 127) // 128
    qp = torch.empty((n, k8), dtype=torch.int32, device=qweight.device)
    sp = torch.empty((tiles, groups, 128), dtype=torch.float32, device=qweight.device)
    zp = torch.empty_like(sp)
    st = stream_handle()
    check(L.b2l_w4a16_pack_qweight(qp.data_ptr(), qweight.data_ptr(), n, k8, qweight.stride(0), qp.stride(0), st))
    check(L.b2l_w4a16_pack_table(sp.data_ptr(), scales.data_ptr(), n, groups, scales.stride(0), scales.stride(1), st))
    check(L.b2l_w4a16_pack_table(zp.data_ptr(), zeros.data_ptr(), n, groups, zeros.stride(0), zeros.stride(1), st))
    foldable = bool(((zeros >= 0) & (zeros <= 1023) & (zeros == zeros.round())).all().item())
    return PackedW4(qp, sp, zp, n, k, int(group_size), foldable)


@torch.no_grad()
def w4a16_matmul_packed(x: torch.Tensor, packed: PackedW4, *, bias: torch.Tensor | None = None,
                        swiglu: bool = False) -> torch.Tensor:
    """:func:`w4a16_matmul` on a :class:`PackedW4` (same result, bit for bit in the integer unpack).

    ``swiglu=True``: the packed rows are interleaved ``(gate_0, up_0, gate_1, up_1, ...)`` (see
    :func:`interleave_gate_up`) and the GEMM's epilogue applies ``silu(gate) * up``; the result is ``[..., N/2]``
    and equals ``swiglu_forward`` of the two separate projections bit for bit."""
    if x.dtype != torch.float16:
        raise ValueError(f"w4a16 activations must be fp16, got {x.dtype}")
    require_cuda(x, bias)
    n, k = packed.n, packed.k
    if x.shape[-1] != k:
        raise ValueError(f"x has {x.shape[-1]} cols but weight expects {k}")
    leading = x.shape[:-1]
    a = x.reshape(-1, k)
    if a.stride(-1) != 1 or a.stride(0) % 8 != 0 or a.data_ptr() % 16 != 0:
        a = a.contiguous()
    m = a.shape[0]
    n_out = n // 2 if swiglu else n
    out = torch.empty((m, n_out), dtype=x.dtype, device=x.device)
    bias_f32 = 0
    if bias is not None:
        if bias.dtype == torch.float32:
            bias_f32 = 1
        elif bias.dtype != torch.float16:
            bias = bias.float()
            bias_f32 = 1
        bias = bias.contiguous()
    L = lib()
    nbytes = L.b2l_quant_gemm_workspace_bytes(_Q_INT4, m, n, k, packed.group_size)
    ws = zeroed_workspace(nbytes, x.device)
    cnt = counters(x.device)
    check(
        L.b2l_w4a16_matmul_packed(
            out.data_ptr(), a.data_ptr(), packed.qweight.data_ptr(), packed.scales.data_ptr(), packed.zeros.data_ptr(),
            ptr(bias), bias_f32, m, n, k, packed.group_size, 1 if packed.zeros_foldable else 0, a.stride(0),
            packed.qweight.stride(0), out.stride(0), ptr(ws), nbytes, cnt.data_ptr(), cnt.numel(), 1 if swiglu else 0,
            stream_handle(),
        )
    )
    return out.reshape(*leading, n_out)


def interleave_gate_up(t: torch.Tensor) -> torch.Tensor:
    """Rows ``[gate_0..gate_{I-1}, up_0..up_{I-1}]`` -> ``[gate_0, up_0, gate_1, up_1, ...]`` (any row-wise tensor of
    the gate|up projection: packed weights, scales, zero points, bias). Load-time only."""
    half = t.shape[0] // 2
    return torch.stack([t[:half], t[half:]], dim=1).reshape(t.shape).contiguous()

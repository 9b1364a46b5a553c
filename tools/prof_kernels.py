"""Tiny driver for ncu captures of single kernels at BASELINE shapes (see profiles/README.md)."""

from __future__ import annotations

import argparse
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lite_llama_b200.kernels as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", required=True)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--n", type=int, default=28672)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--ctx", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--swiglu", action="store_true", help="w4a16 packed: SiLU*up epilogue over interleaved rows")
    ap.add_argument("--reference-layout", action="store_true", help="w4a16: the reference's tensors instead of the load-time repack")
    a = ap.parse_args()
    dev = "cuda"
    if a.op in ("w4a16", "w8a16", "dense", "w8a8"):
        x = torch.randn(a.m, a.k, device=dev, dtype=torch.float16) * 0.5
        if a.op == "w4a16":
            qw = torch.randint(-(2**31), 2**31 - 1, (a.n, a.k // 8), device=dev, dtype=torch.int64).to(torch.int32)
            sc = torch.rand(a.n, a.k // 128, device=dev) * 0.01
            zp = torch.full((a.n, a.k // 128), 8.0, device=dev)
            if a.reference_layout:
                fn = lambda: K.w4a16_matmul(x, qw, sc, zp, group_size=128)
            else:
                pk = K.w4a16_pack(qw, sc, zp, group_size=128)
                fn = lambda: K.w4a16_matmul_packed(x, pk, swiglu=a.swiglu)
        elif a.op == "w8a16":
            qw = torch.randint(0, 120, (a.n, a.k), device=dev, dtype=torch.uint8)
            sc = torch.rand((a.n + 127) // 128, (a.k + 127) // 128, device=dev)
            fn = lambda: K.w8a16_matmul(x, qw, sc, group_n=128, group_k=128)
        elif a.op == "w8a8":
            qw = torch.randint(-127, 127, (a.n, a.k), device=dev, dtype=torch.int8)
            sc = torch.rand(a.n, device=dev) * 0.01
            fn = lambda: K.smoothquant_matmul(x, qw, sc)
        else:
            w = torch.randn(a.n, a.k, device=dev, dtype=torch.float16) * 0.05
            fn = lambda: K.linear(x, w)
    elif a.op == "decode":
        B, hq, hkv, d = 32, 32, 8, 128
        rows = B * a.ctx
        pool = torch.randn(rows, 2 * hkv, d, device=dev, dtype=torch.float16)
        table = torch.arange(rows, dtype=torch.int32, device=dev).view(B, a.ctx)
        q = torch.randn(B, hq, d, device=dev, dtype=torch.float16)
        seq = torch.full((B,), a.ctx, dtype=torch.int32, device=dev)
        req = torch.arange(B, dtype=torch.int32, device=dev)
        fn = lambda: K.flash_decoding(q, pool[:, :hkv], pool[:, hkv:], 1 / math.sqrt(d), table, req, seq, a.ctx)
    elif a.op == "prefill":
        # causal varlen prefill, Llama-3-8B heads, q/k/v as views of one fused-qkv buffer
        B, S, hq, hkv, d = a.m, a.ctx, 32, 8, 128
        T = B * S
        qkv = torch.randn(T, (hq + 2 * hkv) * d, device=dev, dtype=torch.float16) * 0.3
        q = qkv[:, : hq * d].view(T, hq, d)
        k = qkv[:, hq * d:(hq + hkv) * d].view(T, hkv, d)
        v = qkv[:, (hq + hkv) * d:].view(T, hkv, d)
        start = torch.arange(B, device=dev, dtype=torch.int32) * S
        lens = torch.full((B,), S, dtype=torch.int32, device=dev)
        fn = lambda: K.flash_attention2_no_pad(q, k, v, 1.4426950408889634 / math.sqrt(d), start, lens, S)
    elif a.op == "moe":
        # Qwen3-30B-A3B layer, fp8 experts with 128 x 128 block scales, a.m tokens
        E, top_k, I, h, T = 128, 8, 768, 2048, a.m
        hidden = torch.randn(T, h, device=dev, dtype=torch.float16) * 0.5
        w1 = (torch.randn(E, 2 * I, h, device=dev) * 0.5).to(torch.float8_e4m3fn).view(torch.uint8)
        w2 = (torch.randn(E, h, I, device=dev) * 0.5).to(torch.float8_e4m3fn).view(torch.uint8)
        s1 = torch.rand(E, (2 * I + 127) // 128, h // 128, device=dev) * 0.01 + 0.005
        s2 = torch.rand(E, h // 128, (I + 127) // 128, device=dev) * 0.01 + 0.005
        tw, ti = torch.topk(torch.softmax(torch.randn(T, E, device=dev), -1), top_k, dim=-1)
        tw = (tw / tw.sum(-1, keepdim=True)).to(torch.float16)
        ti = ti.to(torch.int32)
        fn = lambda: K.fused_moe(hidden, w1, w2, tw, ti, w1_scale=s1, w2_scale=s2, group_n=128, group_k=128)
    else:
        raise SystemExit(f"unknown op {a.op}")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(a.iters):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    print("done", a.op)


if __name__ == "__main__":
    main()

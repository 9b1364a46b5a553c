"""A/B timing of the W4A16 GEMM at the Llama-3-8B projection shapes under the library's debug flags
(b2l_debug_set_flags; gemm_quant.cu: DBG_*). Per-launch microseconds from a CUDA graph of `reps` launches over
distinct weight tensors (working set >> L2), like bench.py's kernel breakdown. Timing-only flags give wrong results
on purpose; they price an idea before it is built."""

from __future__ import annotations

import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lite_llama_b200.kernels as K  # noqa: E402
from lite_llama_b200 import _lib  # noqa: E402

SHAPES = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}
FLAGS = {"ops8": 65536, "segwait": 131072, "ops8+segwait": 196608, "dbg_inst": 32768, "noX": 4096, "halfring": 8192, "skipall": 144+768, "skipall+noX": 144+768+4096, "skipall+halfring": 144+768+8192, "spin": 1024, "short": 2048, "A+few_mma+no_sttm+spin": 144+768+1024, "A+few_mma+no_sttm+short": 144+768+2048, "no_streamk+spin": 64+1024, "no_streamk+short": 64+2048, "A": 144, "A+few_mma": 144 + 256, "A+no_sttm": 144 + 512, "A+few_mma+no_sttm": 144 + 768, "few_mma": 256, "no_sttm": 512, "skip_drain": 128, "skip_drain+raw": 144, "skip_drain+pair_k4": 160, "base": 0, "raw_operand": 16, "pair_k4": 32, "no_streamk": 64, "no_streamk+pair_k4": 96, "no_streamk+raw": 80}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--shapes", default="qkv,o,gate_up,down")
    ap.add_argument("--flags", default="base,no_streamk,pair_k4,raw_operand,no_streamk+pair_k4,no_streamk+raw")
    a = ap.parse_args()
    lib = _lib.load()
    lib.b2l_debug_set_flags.argtypes = [ctypes.c_int]
    lib.b2l_debug_set_flags.restype = ctypes.c_int
    dev = "cuda"
    ap2 = [x for x in a.shapes.split(",")]
    for name, (n, k) in SHAPES.items():
        if name not in ap2:
            continue
        x = torch.randn(a.m, k, device=dev, dtype=torch.float16) * 0.5
        ws = []
        for _ in range(a.reps):
            qw = torch.randint(-(2**31), 2**31 - 1, (n, k // 8), device=dev, dtype=torch.int64).to(torch.int32)
            sc = torch.rand(n, k // 128, device=dev) * 0.01
            zp = torch.full((n, k // 128), 8.0, device=dev)
            ws.append((qw, sc, zp))
        nbytes = n * k // 2 + n * (k // 128) * 8
        row = {"shape": name, "m": a.m, "n": n, "k": k}
        packed = [K.w4a16_pack(qw, sc, zp, group_size=128) for qw, sc, zp in ws]
        for fname in a.flags.split(","):
            use_packed = fname.startswith("packed")
            lib.b2l_debug_set_flags(FLAGS[fname[7:] if fname.startswith("packed+") else ("base" if use_packed else fname)])

            def fn():
                if use_packed:
                    for pk in packed:
                        K.w4a16_matmul_packed(x, pk)
                else:
                    for qw, sc, zp in ws:
                        K.w4a16_matmul(x, qw, sc, zp, group_size=128)

            fn()
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            ts = []
            for _ in range(7):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g.replay()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3 / a.reps)
            ts.sort()
            us = ts[len(ts) // 2]
            row[fname] = round(us, 2)
            row[fname + "_gbs"] = round(nbytes / us / 1e3, 0)
            del g
        lib.b2l_debug_set_flags(0)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

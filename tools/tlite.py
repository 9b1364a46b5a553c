"""Per-CTA clock stamps of the PRODUCT W4A16 kernel (library built with -DB2L_TLITE, tools/build_variants.py tlite):
where a CTA's lifetime goes (set-up, first bytes, pipeline fill, steady state, tail). Not part of the product path."""
from __future__ import annotations

import argparse
import ctypes
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lite_llama_b200.kernels as K  # noqa: E402
from lite_llama_b200 import _lib  # noqa: E402

NAMES = {23: "first instruction", 24: "producer lane: reached its set-up", 25: "producer lane: weight-ring barriers initialised", 26: "producer lane: init fences done", 0: "entry", 19: "tables requested (producer lane)", 1: "first weight requests out", 2: "set-up barrier passed",
         3: "x producer: griddepcontrol.wait returned", 8: "dequant g0: first weights in smem", 9: "dequant g0: first block stored",
         4: "mma0: first activations landed", 5: "mma0: first block issued", 12: "epilogue: first segment complete",
         10: "dequant g0: last block stored", 11: "dequant g1: last block stored", 6: "mma0: last block issued",
         7: "mma1: last block issued", 13: "epilogue: last segment drained", 14: "epilogue: tile stored / ticket taken",
         15: "final barrier passed", 28: "fix-up: own partial read, polling starts", 30: "fix-up: first batch of loads back", 27: "fix-up: first tile summed", 20: "fix-up done (exit)"}
ORDER = [23, 0, 24, 25, 26, 19, 1, 2, 3, 8, 9, 4, 5, 12, 10, 11, 6, 7, 13, 14, 15, 28, 30, 27, 20]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="o:4096x4096,qkv:6144x4096,down:4096x14336,gate_up:28672x4096,one_rb:18944x256,four_rb:18944x1024")
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--chain", type=int, default=6)
    a = ap.parse_args()
    lib = _lib.load()
    lib.b2l_debug_set_timeline.argtypes = [ctypes.c_void_p]
    lib.b2l_debug_set_timeline.restype = ctypes.c_int
    dev = "cuda"
    for spec in a.shapes.split(","):
        name, nk = spec.split(":")
        n, k = map(int, nk.split("x"))
        x = torch.randn(a.m, k, device=dev, dtype=torch.float16) * 0.5
        packed = []
        for _ in range(a.chain):
            qw = torch.randint(-(2**31), 2**31 - 1, (n, k // 8), device=dev, dtype=torch.int64).to(torch.int32)
            sc = torch.rand(n, k // 128, device=dev) * 0.01
            zp = torch.full((n, k // 128), 8.0, device=dev)
            packed.append(K.w4a16_pack(qw, sc, zp, group_size=128))
        for pk in packed:
            K.w4a16_matmul_packed(x, pk)
        torch.cuda.synchronize()
        buf = torch.zeros(1024 * 32, dtype=torch.int64, device=dev)
        lib.b2l_debug_set_timeline(buf.data_ptr())
        for pk in packed:
            K.w4a16_matmul_packed(x, pk)
        torch.cuda.synchronize()
        lib.b2l_debug_set_timeline(None)
        t = buf.cpu().view(1024, 32)
        ctas = [i for i in range(1024) if t[i, 0] != 0]
        print(f"== {name} m={a.m} n={n} k={k}: {len(ctas)} CTAs, raw blocks per CTA {int(t[ctas[0], 22])}..{int(max(t[c, 22] for c in ctas))}; "
              f"CTAs that reduced a shared tile: {sum(1 for c in ctas if t[c, 21] > 0)}")
        g0 = min(int(t[c, 16]) for c in ctas)
        life = [int(t[c, 17] - t[c, 16]) for c in ctas]
        print(f"   wall clock (globaltimer): first entry -> last exit {max(int(t[c, 17]) for c in ctas) - g0} ns; "
              f"CTA lifetime min/median/max {min(life)}/{int(statistics.median(life))}/{max(life)} ns; entry spread {max(int(t[c, 16]) for c in ctas) - g0} ns")
        print("   stamp (ns after the CTA's entry)                      min   median      max")
        for idx in ORDER:
            vals = []
            for c in ctas:
                if t[c, idx] == 0 and idx != 0:
                    continue
                cyc = int(t[c, idx] - t[c, 0])
                tot_cyc = int(t[c, 20] - t[c, 0])
                ns = cyc * (int(t[c, 17] - t[c, 16]) / tot_cyc) if tot_cyc > 0 else 0.0
                vals.append(ns)
            if vals:
                print(f"   {NAMES[idx]:50s} {min(vals):8.0f} {statistics.median(vals):8.0f} {max(vals):8.0f}   (n={len(vals)})")
        worst = sorted(ctas, key=lambda c: -int(t[c, 17] - t[c, 16]))[:4]
        print("   the four longest-lived CTAs (ns after their entry):")
        for c in worst:
            tot_cyc = int(t[c, 20] - t[c, 0])
            sc = int(t[c, 17] - t[c, 16]) / tot_cyc
            row = " ".join(f"{NAMES[i].split(':')[0][:10]}#{i}={int((t[c, i] - t[c, 0]) * sc)}" for i in ORDER if t[c, i] != 0 and i not in (0, 23, 24, 25, 26))
            print(f"     cta {c} rb={int(t[c, 22])} nfix={int(t[c, 21])} failed polls={int(t[c, 29])}: {row}")


if __name__ == "__main__":
    main()

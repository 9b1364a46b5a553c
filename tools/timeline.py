"""Device timeline of one CTA of the quantised GEMM (debug hook b2l_debug_set_timeline): prints where each
role spends its cycles per raw block (256 int4 / 128 8-bit K values). Used to find pipeline stalls; not part of the product path."""

from __future__ import annotations

import argparse
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lite_llama_b200.kernels as K  # noqa: E402
from lite_llama_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="w4a16")
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--n", type=int, default=28672)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--hot", type=int, default=0, help="launches issued right before the stamped one (sustained-load clocks)")
    ap.add_argument("--brief", action="store_true")
    ap.add_argument("--packed", action="store_true", help="kernel-native W4A16 layout (w4a16_pack)")
    a = ap.parse_args()
    lib = _lib.load()
    lib.b2l_debug_set_timeline.argtypes = [ctypes.c_void_p]
    lib.b2l_debug_set_timeline.restype = ctypes.c_int
    lib.b2l_debug_set_flags.argtypes = [ctypes.c_int]
    lib.b2l_debug_set_flags.restype = ctypes.c_int
    dev = "cuda"
    x = torch.randn(a.m, a.k, device=dev, dtype=torch.float16) * 0.5
    if a.op == "w4a16":
        sets = []
        for _ in range(4 if a.hot else 1):   # hot mode cycles through 4 weight sets so that no launch finds its weights in L2
            qw = torch.randint(-(2**31), 2**31 - 1, (a.n, a.k // 8), device=dev, dtype=torch.int64).to(torch.int32)
            sc = torch.rand(a.n, a.k // 128, device=dev) * 0.01
            zp = torch.full((a.n, a.k // 128), 8.0, device=dev)
            sets.append((qw, sc, zp))
        calls = [0]

        packed = [K.w4a16_pack(*t, group_size=128) for t in sets] if a.packed else None

        def fn():
            i = calls[0] % len(sets)
            calls[0] += 1
            if packed is not None:
                return K.w4a16_matmul_packed(x, packed[i])
            qw, sc, zp = sets[i]
            return K.w4a16_matmul(x, qw, sc, zp, group_size=128)
    else:
        qw = torch.randint(0, 120, (a.n, a.k), device=dev, dtype=torch.uint8)
        sc = torch.rand((a.n + 127) // 128, (a.k + 127) // 128, device=dev)
        fn = lambda: K.w8a16_matmul(x, qw, sc, group_n=128, group_k=128)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tl = torch.zeros(16 * 256 + 3 * 1024, dtype=torch.int64, device=dev)
    lib.b2l_debug_set_flags(a.flags)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush.zero_()
    for _ in range(a.hot):
        fn()
    lib.b2l_debug_set_timeline(tl.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.b2l_debug_set_timeline(None)
    lib.b2l_debug_set_flags(0)
    cta = tl.cpu()[4096:].view(1024, 3)
    steps = tl.cpu()[2048:4096].view(128, 16)
    t = tl.cpu()[:2048].view(8, 256)
    live = cta[:, 0] > 0
    if bool(live.any()):
        st, en, sm = cta[live, 0], cta[live, 1], cta[live, 2]
        g0 = int(st.min())
        dur = (en - st).float()
        print(f"== {a.op} m={a.m} n={a.n} k={a.k} flags={a.flags} hot={a.hot}: {int(live.sum())} CTAs, kernel wall {int(en.max()) - g0} ns; "
              f"CTA start spread {int(st.max()) - g0} ns; CTA duration min/median/max {int(dur.min())}/{int(dur.median())}/{int(dur.max())} ns; "
              f"last exit by CTA {int(en.argmax())} on SM {int(sm[en.argmax()])}")
        m6, m0_ = int(tl.cpu()[3072 + 6]), int(tl.cpu()[3072])
        if m6 > 0 and int(cta[0, 1]) > int(cta[0, 0]):
            print(f"   CTA 0: {m6 - m0_} cycles in {int(cta[0,1]) - int(cta[0,0])} ns -> SM clock {(m6 - m0_) / (int(cta[0,1]) - int(cta[0,0])):.3f} GHz")
        order = torch.argsort(en - g0)
        print("   slowest exits (cta: start,end ns):", [(int(i), int(st[i]) - g0, int(en[i]) - g0) for i in order[-5:]])
        print("   fastest exits:", [(int(i), int(st[i]) - g0, int(en[i]) - g0) for i in order[:3]])
    if a.brief:
        return
    names = ["PROD", "MMA_READY", "MMA_ISSUED", "DQ_START", "DQ_DONE", "EPI_START", "EPI_DONE"]
    t0 = int(t[t > 0].min())
    print(f"== {a.op} m={a.m} n={a.n} k={a.k}: total cycles {int(t[t>0].max()) - t0}")
    misc = tl.cpu()[3072:3072 + 11]
    m0 = int(misc[0])
    names_misc = ["kernel entry", "setup sync done", "epi: tables landed", "epi: zero check + bar4 done",
                  "dequant: past bar4", "epi: last segment drained", "exit sync", "epi: table cp.async issued",
                  "producer: mbarriers initialised", "mma0: TMEM allocated", "warp 0: reached setup sync"]
    print("start-up / shutdown stamps (cycles from kernel entry; first TL stamp at %d):" % (t0 - m0))
    for nm, v in zip(names_misc, misc.tolist()):
        print(f"  {nm:28s} {int(v) - m0 if v > 0 else -1}")
    print("first events (cycles from first stamp):")
    for r, nm in enumerate(names):
        row = [int(v) - t0 for v in t[r][:6] if v > 0]
        print(f"  {nm:10s} {row}")
    rel = lambda r, i: (int(t[r, i]) - t0) if int(t[r, i]) > 0 else -1
    print("raw-block stamps (cycles from first stamp; -1 = not stamped, e.g. blocks issued by the second MMA warp):")
    print("  rb  prod_issue | mma_top mma_ready mma_issued | dq_start dq_done")
    for i in range(0, min(int(max((t[3] > 0).sum(), (t[1] > 0).sum()) * 2), 32)):
        print(f"  {i:3d} {rel(0,i):10d} | {rel(7,i):7d} {rel(1,i):9d} {rel(2,i):9d} | {rel(3,i):8d} {rel(4,i):8d}")
    nseg = int((t[5] > 0).sum())
    print("segment stamps:  s  epi_start epi_done  (dur)")
    for s_ in range(0, min(nseg, 24)):
        print(f"  {s_:3d} {rel(5,s_):9d} {rel(6,s_):9d}  ({rel(6,s_)-rel(5,s_)})")
    print("MMA fast path per segment: acc_empty wait done -> issue done (cycles from first stamp)")
    flat = steps.reshape(-1)
    for sg in range(8, 20):
        a0, a1 = int(flat[2 * sg]), int(flat[2 * sg + 1])
        if a0 > 0:
            print(f"  seg {sg}: {a0 - t0} -> {a1 - t0}  ({a1 - a0})")
    sub = tl.cpu()[3072 + 32:3072 + 32 + 96 * 5].view(96, 5)
    if int((sub > 0).sum()) > 0:
        print("dequant warp 0, per 64-element sub-block (cycles): top->stage free | convert | wait::st+arrive | st issue | gap to next top")
        for i in range(96):
            if int(sub[i, 0]) <= 0:
                continue
            nxt = [int(sub[j, 0]) for j in range(i + 1, 96) if int(sub[j, 0]) > 0]
            gap = (nxt[0] - int(sub[i, 4])) if nxt else -1
            print(f"  gsub {i:3d} @{int(sub[i,0]) - t0:7d}: {int(sub[i,1]-sub[i,0]):5d} | {int(sub[i,2]-sub[i,1]):5d} | {int(sub[i,3]-sub[i,2]):5d} | {int(sub[i,4]-sub[i,3]):5d} | {gap:5d}")
    nrb = int((t[0] > 0).sum())
    


if __name__ == "__main__":
    main()

"""Per-kernel SASS instruction summary of libb200lite.so's objects: the mnemonics that prove which hardware path a
kernel uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, legacy
tensor path -> HMMA/IMMA, cp.async -> LDGSTS).  Runs on the CPU build box (cuobjdump only).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMASTG", "HMMA", "IMMA", "LDGSTS", "MUFU.EX2",
        "SYNCS", "ATOM", "RED", "LDSM", "SHFL"]


def main():
    objs = sorted((ROOT / "lite_llama_b200" / "csrc" / "build").glob("*.o"))
    print("# SASS mnemonic counts per kernel (cuobjdump -sass on the sm_100a objects linked into libb200lite.so)")
    print("# columns: " + " ".join(KEYS) + " | total instructions")
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        cur = None
        counts: dict[str, Counter] = defaultdict(Counter)
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and cur:
                op = m.group(1)
                counts[cur]["_total"] += 1
                for k in KEYS:
                    if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                        counts[cur][k] += 1
        if not counts:
            continue
        print(f"\n## {obj.name}")
        names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
        for mangled, nice in zip(counts, names):
            c = counts[mangled]
            short = re.sub(r"\(.*", "", nice)[:110]
            cells = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
            print(f"{short}: {cells or '-'} | {c['_total']}")


if __name__ == "__main__":
    sys.exit(main())

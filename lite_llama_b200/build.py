"""Build the in-tree C-ABI shared library ``libb200lite.so`` with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the GPU-less build container; the ``.so``
then travels with the repo snapshot to the B200 box. No torch headers are involved: the
library only depends on the (statically linked) CUDA runtime.

Usage:
    python -m lite_llama_b200.build            # incremental
    python -m lite_llama_b200.build --force
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
OUT = Path(__file__).resolve().parent / "libb200lite.so"
OBJ_DIR = CSRC / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, force: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    stamp_file = OBJ_DIR / (src.stem + ".stamp")
    stamp = _stamp(src)
    if not force and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    OBJ_DIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(lambda s: _compile(s, force), sources))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not OUT.exists() or OUT.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-o", str(OUT), *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[b200lite] linked {OUT} from {len(objs)} objects")
    elif verbose:
        print(f"[b200lite] {OUT.name} is up to date")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)

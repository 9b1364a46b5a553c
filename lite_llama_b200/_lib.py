"""ctypes loader for ``libb200lite.so``.

Prototypes are read from ``include/b200_lite_kernels.h`` so the Python binding cannot drift from
the declared C ABI. The library is required: there is deliberately no fallback path.
"""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libb200lite.so"
HEADER_PATH = _PKG.parent / "include" / "b200_lite_kernels.h"

_CTYPES = {
    "void*": ctypes.c_void_p,
    "const void*": ctypes.c_void_p,
    "int64_t": ctypes.c_int64,
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "int*": ctypes.POINTER(ctypes.c_int),
    "const char*": ctypes.c_char_p,
}

_PROTO = re.compile(r"^\s*(const char\*|int64_t|int)\s+(b2l_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def parse_header(path: Path = HEADER_PATH) -> dict[str, tuple[str, list[str]]]:
    """Return ``{symbol: (return_type, [arg types])}`` for every prototype in the header."""
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    out: dict[str, tuple[str, list[str]]] = {}
    for ret, name, args in _PROTO.findall(text):
        arg_types: list[str] = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"^(const\s+)?(void|char|int64_t|int|float)\s*(\*?)\s*\w*$", a)
                if not m:
                    raise RuntimeError(f"cannot parse argument {a!r} of {name}")
                t = f"{'const ' if m.group(1) else ''}{m.group(2)}{m.group(3)}"
                arg_types.append(t)
        out[name] = (ret, arg_types)
    return out


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library with typed prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LibraryMissing(
            f"{LIB_PATH} not found. The B200 kernel library is required (no fallback); build it with "
            "`python -m lite_llama_b200.build` (nvcc, sm_100a)."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (ret, args) in parse_header().items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # header declares a symbol the library does not export
            raise LibraryMissing(f"{LIB_PATH.name} does not export {name}; rebuild the library") from e
        fn.restype = _CTYPES[ret]
        fn.argtypes = [_CTYPES[a] for a in args]
    _lib = lib
    return lib


def _fnv1a(name: str) -> int:
    h = 2166136261
    for c in name.encode():
        h = ((h ^ c) * 16777619) & 0xFFFFFFFF
    return h


def hang_report() -> dict | None:
    """Decode the device-side hang watchdog's record (csrc/common.cuh): which barrier wait of which kernel
    source gave up. ``None`` when no kernel of the library has trapped. Works after the CUDA context died."""
    lib = load()
    buf = (ctypes.c_int * 16)()
    lib.b2l_hang_report(buf)
    w = [v & 0xFFFFFFFF for v in buf]
    if w[0] != 0xB2D00DED:
        return None
    names = {_fnv1a(f.name): f.name for f in (_PKG / "csrc").glob("*.cu*")}
    return {
        "file": names.get(w[1], hex(w[1])), "line": w[2], "block": (w[3], w[4], w[5]), "thread": w[6],
        "warp": w[6] // 32, "barrier_smem_addr": hex(w[7]), "parity": w[8], "waited_ms": w[9],
        "grid": (w[10], w[11], w[12]), "block_dim": w[13],
    }


def last_error() -> str:
    return load().b2l_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Raise ``ValueError`` (the reference's error type for bad shapes/dtypes) on a non-zero return."""
    if rc != 0:
        raise ValueError(f"b200lite error {rc}: {last_error()}")

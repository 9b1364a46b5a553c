"""A/B builds of the quantised GEMM: compiles csrc/gemm_quant.cu with -D knobs and links each against the other
objects of the regular build into tools/variants/lib_<name>.so (git-ignored, travels to the GPU box).
tools/run_variants.sh swaps each one in for libb200lite.so on the box and runs the probe. Not part of the product."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from lite_llama_b200 import build as B  # noqa: E402

VARIANTS = {
    "base": [],
    "short": ["-DB2L_QWAIT_MODE=2"],
    "spin_lat": ["-DB2L_QWAIT_LAT_MODE=1"],
    "short_lat": ["-DB2L_QWAIT_LAT_MODE=2"],
    "order1": ["-DB2L_DQ_ORDER=1"],
    "order1_spin_lat": ["-DB2L_DQ_ORDER=1", "-DB2L_QWAIT_LAT_MODE=1"],
    "order1_short": ["-DB2L_DQ_ORDER=1", "-DB2L_QWAIT_MODE=2"],
    "tlite": ["-DB2L_TLITE=1"],
    "cap32": ["-DB2L_QUANT_BT_CAP=32"],
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    B.build(force=False, verbose=True)
    out_dir = ROOT / "tools" / "variants"
    out_dir.mkdir(exist_ok=True)
    others = [str(o) for o in sorted(B.OBJ_DIR.glob("*.o")) if o.stem != "gemm_quant"]
    procs = []
    for name in names:
        obj = out_dir / f"gemm_quant_{name}.o"
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *VARIANTS[name], "-c", str(B.CSRC / "gemm_quant.cu"), "-o", str(obj)]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("variant compile failed")
    for name in names:
        obj = out_dir / f"gemm_quant_{name}.o"
        lib = out_dir / f"lib_{name}.so"
        subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-o", str(lib), str(obj), *others], check=True)
        print("built", lib)


if __name__ == "__main__":
    main()

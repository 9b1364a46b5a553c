"""Pretty-print gpurun_out/microbench.jsonl and bench_short.json (local helper)."""
import json, sys
from pathlib import Path
out = Path(__file__).resolve().parent.parent / "gpurun_out"
mb = out / "microbench.jsonl"
if mb.exists() and "--no-micro" not in sys.argv:
    for l in mb.read_text().splitlines():
        r = json.loads(l)
        if "error" in r:
            print(r); continue
        if r["op"] == "flash_decoding":
            print(f"decode B={r['B']} ctx={r['ctx']} scat={r['scattered']} {r['ms']*1000:.1f}us {r['gbs']:.0f}GB/s frac={r['frac_hbm']:.2f}")
        elif r["m"] <= 32:
            print(f"{r['op']:6s} m={r['m']:4d} n={r['n']:6d} k={r['k']:5d} {r['ms']*1000:8.1f}us {r['gbs']:7.0f}GB/s fracHBM={r.get('frac_hbm',0):.2f}")
        else:
            print(f"{r['op']:6s} m={r['m']:4d} n={r['n']:6d} k={r['k']:5d} {r['ms']*1000:8.1f}us {r['tflops']:.1f}TF fracTC={r.get('frac_tensor',0):.3f}")
name = next((a for a in sys.argv[1:] if a.endswith(".json")), "bench_short.json")
b = out / name
if b.exists():
    d = json.loads(b.read_text().strip().splitlines()[-1])
    for k in ["value", "ms_per_step", "e2e", "gpu_launches", "clocks", "step_roofline", "prefill", "roofline", "cpu_baseline"]:
        print(k, d.get(k))
    for r in d.get("kernels", []):
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k in ("kernel", "us_per_launch", "achieved_gbs", "frac_hbm", "us_per_step")})
    if d.get("kernels_error"): print(d["kernels_error"])

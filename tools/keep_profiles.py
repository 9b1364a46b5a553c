"""Condense what a GPU pass left in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/keep_profiles.py r1b        # prefix for this round/pass

* launches.csv (ncu --metrics gpu__time_duration.sum over `bench.py --steps 2 --warmup 1`): one decode step's
  launch list aggregated per kernel, with each kernel's SHARE of the step (per-launch times under ncu are
  cold-cache and serialised, so only shares are comparable with bench.py's live numbers).
* *.ncu-rep (--set full captures): the handful of metrics DESIGN.md quotes (duration, DRAM bytes, throughput %,
  pipe utilisation, issue-slot use, stall reasons) per capture.
* bench*.json: copied verbatim.
"""

from __future__ import annotations

import csv
import json
import re
import shutil
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

OURS = ["gemm_quant_kernel", "gemm_swapab_kernel", "moe_gemm_kernel", "flash_decode_split_kernel",
        "flash_decode_combine_kernel", "flash_prefill_kernel", "skip_rmsnorm_block_kernel", "skip_rmsnorm_warp_reg_kernel",
        "skip_rmsnorm_warp_kernel", "rope_kernel", "rope_kv_kernel", "swiglu_kernel", "update_kv_buffer_kernel",
        "update_kv_index_kernel", "quantize_act_int8_kernel", "moe_sum_kernel", "moe_align_kernel", "moe_route_kernel",
        "argmax_rows_kernel", "embed_rope_gather_kernel", "rope_kv_write_kernel", "fa_tc_kernel", "fa_tc_one_tile_kernel",
        "tp_allreduce_rmsnorm_kernel"]


def kname(full: str) -> str:
    for k in OURS:
        if k in full:
            m = re.search(k + r"<([^>]*)>", full)
            return k + (f"<{m.group(1)}>" if m and len(m.group(1)) < 40 else "")
    m = re.search(r"(\w+_kernel|\w+Kernel)\b", full)
    return "torch:" + (m.group(1) if m else full[:40])


def launches(prefix: str) -> None:
    src = OUT / "launches.csv"
    if not src.exists():
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit() and r[-3] == "gpu__time_duration.sum"]
    seq = []
    for r in rows:
        v = float(r[-1])
        unit = r[-2]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        seq.append((kname(r[4]), r[8], r[7], us))
    heads = [i for i, s in enumerate(seq) if s[0].startswith("gemm_swapab_kernel")]
    # decode step = launches between two consecutive lm_head GEMMs; take the last window that looks like a full step
    window = None
    for a, b in zip(reversed(heads[:-1]), reversed(heads[1:])):
        if b - a >= 40 and sum(1 for s in seq[a + 1:b + 1] if "flash_decode_split" in s[0]) > 0:
            window = (a + 1, b + 1)
            break
    lines = [f"# source: gpurun_out/launches.csv ({len(seq)} launches in the whole bench.py run)"]
    if window is None:
        lines.append("# no decode-step window found")
    else:
        step = seq[window[0]:window[1]]
        tot = sum(s[3] for s in step)
        agg: "OrderedDict[str, list]" = OrderedDict()
        for n, g, b, us in step:
            e = agg.setdefault(n, [0, 0.0, set()])
            e[0] += 1
            e[1] += us
            e[2].add(f"{g}x{b}")
        ours = sum(e[0] for n, e in agg.items() if not n.startswith("torch:"))
        lines.append(f"# one decode step = launches [{window[0]}, {window[1]}) : {len(step)} kernels ({ours} of this library), "
                     f"sum of per-launch durations {tot:.1f} us (cold-cache, serialised under ncu: compare SHARES only)")
        lines.append(f"{'kernel':58s} {'launches':>8s} {'us/launch':>10s} {'us/step':>10s} {'share':>7s}  grid x block")
        for n, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{n:58s} {e[0]:8d} {e[1] / e[0]:10.2f} {e[1]:10.1f} {100 * e[1] / tot:6.1f}%  {' '.join(sorted(e[2]))[:60]}")
    (PROF / f"{prefix}_launches_decode_step.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def ncu_reports(prefix: str) -> None:
    for rep in sorted(OUT.glob("*.ncu-rep")):
        res = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(res.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        lines = [f"# ncu --set full --clock-control none, report {rep.name} (kept in gpurun_out/, not tracked)"]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            lines.append(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}")
            for k in KEYS:
                if k in d:
                    lines.append(f"  {k:86s} {d[k][1]:>16s} {d[k][0]}")
            for k, (u, v) in d.items():
                if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                    try:
                        if float(v) >= 0.2:
                            lines.append(f"  {k:86s} {float(v):16.2f}")
                    except ValueError:
                        pass
        (PROF / f"{prefix}_ncu_{rep.stem}.txt").write_text("\n".join(lines) + "\n")
        print(f"wrote profiles/{prefix}_ncu_{rep.stem}.txt")


def main() -> None:
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r1"
    PROF.mkdir(exist_ok=True)
    launches(prefix)
    ncu_reports(prefix)
    for f in OUT.glob("bench*.json"):
        txt = f.read_text().strip()
        if txt:
            try:
                json.loads(txt.splitlines()[-1])
            except json.JSONDecodeError:
                continue
            shutil.copy(f, PROF / f"{prefix}_{f.name}")
    for f in ("timeline.txt", "microbench.jsonl"):
        if (OUT / f).exists() and (OUT / f).stat().st_size:
            shutil.copy(OUT / f, PROF / f"{prefix}_{f}")


if __name__ == "__main__":
    main()

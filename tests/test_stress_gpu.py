"""Graph-replay stress of the captured decode step at Llama-3-8B geometry (round-1 VERDICT item 1: intermittent
device hangs that the ~12-replay model tests could not see).  Each case runs tools/stress_decode.py in its own
process (a trapped kernel kills the CUDA context): 4 real-shape W4A16 layers + lm_head, the whole greedy step as one
CUDA graph, replayed 2000 times with the token stream compared against a first pass.  The device-side hang
watchdog (csrc/common.cuh) turns a dead-locked pipeline into an error within seconds and names the barrier."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _run(args: list[str], timeout: int = 240) -> dict:
    cmd = [sys.executable, str(ROOT / "tools" / "stress_decode.py"), *args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, f"no result line (rc={res.returncode})\nstdout:\n{res.stdout[-2000:]}\nstderr:\n{res.stderr[-2000:]}"
    out = json.loads(lines[-1])
    assert out.get("ok"), f"stress run failed: {out}\nstderr:\n{res.stderr[-1500:]}"
    return out


@pytest.mark.parametrize("batch", [1, 2, 32])
@pytest.mark.parametrize("pdl", [1, 0])
def test_decode_graph_replay_2000(batch, pdl):
    out = _run(["--batch", str(batch), "--layers", "4", "--replays", "2000", "--pdl", str(pdl)])
    assert out["mismatching_passes"] == 0


def test_two_processes_share_the_gpu():
    """Two independent replicas at once (different seeds, like bench.py's ranks)."""
    cmds = [[sys.executable, str(ROOT / "tools" / "stress_decode.py"), "--batch", "32", "--layers", "4", "--replays",
             "1000", "--seed", str(s)] for s in (0, 1)]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for c in cmds]
    outs = []
    for p in procs:
        try:
            so, se = p.communicate(timeout=400)
        except subprocess.TimeoutExpired:
            p.kill()
            so, se = p.communicate()
        outs.append((p.returncode, so, se))
    for rc, so, se in outs:
        lines = [ln for ln in so.splitlines() if ln.startswith("{")]
        assert lines, f"no result line (rc={rc})\n{so[-1500:]}\n{se[-1500:]}"
        res = json.loads(lines[-1])
        assert res.get("ok"), f"{res}\n{se[-1500:]}"


def test_full_depth_batch1_smoke():
    """All 32 layers at batch 1 (the configuration that stalled in round 1), 300 replays."""
    _run(["--batch", "1", "--layers", "32", "--replays", "300"], timeout=400)

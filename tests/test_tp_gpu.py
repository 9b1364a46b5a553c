"""Tensor parallelism on real GPUs (needs >= 2 devices; skipped on a single-GPU box): tools/tp_check.py under
torchrun -- NVLink peer-memory all-reduce vs the fp32 sum and vs NCCL, the fused all-reduce + skip_rmsnorm, decoder
logits TP=2 vs TP=1, and the captured greedy loop with the exchange inside the CUDA graph."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensor_parallel_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(ROOT / "tools" / "tp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    lines = [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 2, f"rc={res.returncode}\n{res.stdout[-3000:]}\n{res.stderr[-3000:]}"
    for ln in lines:
        assert ln["ok"], ln

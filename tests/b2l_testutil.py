"""Helpers shared by the test modules (golden fixture loading, error metrics)."""

from __future__ import annotations

from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def load_golden(name: str) -> dict:
    return torch.load(GOLDEN / f"{name}.pt", map_location="cpu", weights_only=True)


def max_abs_diff(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.float() - b.float()).abs().max().item()


def to_cuda(obj):
    """Recursively move the tensors of a golden case to the GPU."""
    if isinstance(obj, torch.Tensor):
        return obj.cuda()
    if isinstance(obj, dict):
        return {k: to_cuda(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_cuda(v) for v in obj)
    return obj

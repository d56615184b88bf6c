"""Optional per-stage CUDA-event timing (off by default; bench.py turns it on for the
roofline block).  Events are recorded on the current stream around a stage, so the
durations are device times of exactly the kernels of that stage."""
from __future__ import annotations

from collections import defaultdict
from contextlib import contextmanager
from typing import Dict, List

import torch

_enabled = False
_events: Dict[str, List] = defaultdict(list)


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag


def reset() -> None:
    _events.clear()


@contextmanager
def stage(name: str):
    if not _enabled:
        yield
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    try:
        yield
    finally:
        e.record()
        _events[name].append((s, e))


def summary() -> Dict[str, Dict[str, float]]:
    """name -> {count, total_ms, mean_ms}; call after torch.cuda.synchronize()."""
    out = {}
    for name, evs in _events.items():
        ms = [s.elapsed_time(e) for s, e in evs]
        out[name] = dict(count=len(ms), total_ms=float(sum(ms)), mean_ms=float(sum(ms) / max(len(ms), 1)))
    return out

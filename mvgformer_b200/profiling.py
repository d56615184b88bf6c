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
_notes: Dict[str, List] = defaultdict(list)


def enabled() -> bool:
    return _enabled


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag


def reset() -> None:
    _events.clear()
    _notes.clear()


_counters: Dict[str, int] = defaultdict(int)


def count(name: str, n: int = 1) -> None:
    """Host-side event counter (always on): e.g. weight re-packs, which must stay at one per
    module for the life of a model."""
    _counters[name] += n


def counters() -> Dict[str, int]:
    return dict(_counters)


def note(name: str, value) -> None:
    """Keeps a (cloned) device scalar for later inspection, e.g. the number of in-view items the
    gather processed; no synchronisation happens here.  `value` may be a zero-argument callable:
    it is only evaluated when profiling is enabled, so a disabled note enqueues nothing."""
    if _enabled:
        if callable(value) and not isinstance(value, torch.Tensor):
            value = value()
        _notes[name].append(value.detach().clone() if callable(getattr(value, "detach", None)) else value)


def notes() -> Dict[str, List[float]]:
    """name -> list of recorded values; call after torch.cuda.synchronize()."""
    return {k: [float(t) for t in v] for k, v in _notes.items()}


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

"""Process-wide engines for the facades (one HotPathEngine per device, grown on demand)."""

from __future__ import annotations

import threading
from contextlib import contextmanager

import torch

from .engine import HotPathEngine

_lock = threading.RLock()
_engines: dict[int, HotPathEngine] = {}


def resolve_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("slaf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"slaf_b200 computes on CUDA devices only, got {dev}")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def engine_for(n_rows: int, n_cells: int | None = None, device=None) -> HotPathEngine:
    """Shared engine of `device` able to take a batch of n_rows rows (recreated larger if needed)."""
    dev = resolve_device(device)
    need_rows = max(int(n_rows), 1 << 20)
    need_cells = max(int(n_cells if n_cells is not None else n_rows), 1 << 12)
    with _lock:
        eng = _engines.get(dev.index)
        if eng is None or eng.max_rows < need_rows or eng.max_cells < min(need_cells, need_rows):
            if eng is not None:
                eng.close()
            rows = max(need_rows, (eng.max_rows * 2) if eng else 0)
            cells = min(rows, max(need_cells, (eng.max_cells * 2) if eng else 0))
            eng = HotPathEngine(dev, max_rows=rows, max_cells=cells)
            _engines[dev.index] = eng
        return eng


@contextmanager
def session(n_rows: int, n_cells: int | None = None, device=None):
    """The shared engine of `device`, held exclusively for the duration of one facade call: a context is single-threaded
    by contract and may be replaced by a larger one between calls, so facade calls from different threads take turns."""
    with _lock:
        yield engine_for(n_rows, n_cells, device)


def shutdown() -> None:
    with _lock:
        for eng in _engines.values():
            eng.close()
        _engines.clear()


import atexit  # noqa: E402

atexit.register(shutdown)

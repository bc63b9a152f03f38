"""Per-call CUDA-event timing of the C-ABI entry points (used by bench.py for the roofline rows).

When a ``KernelTimer`` is active, every call into libofr_b200.so is bracketed by two CUDA events
recorded on the stream the kernels are launched on (torch's current stream), tagged with the
call's algorithmic bytes and FLOPs (SURVEY.md section 8d formulas).  Nothing is synchronised until
``summary()``.  With no active timer the cost is one global lookup per call.
"""
import contextlib
from collections import OrderedDict

import torch

_active = None


class KernelTimer:
    def __init__(self):
        self.records = []

    def __enter__(self):
        global _active
        self._prev, _active = _active, self
        return self

    def __exit__(self, *exc):
        global _active
        _active = self._prev
        return False

    def summary(self):
        """{label: dict(calls, total_ms, avg_us, bytes, flops, gbs, tflops)} after a device synchronise."""
        torch.cuda.synchronize()
        out = OrderedDict()
        for label, nbytes, flops, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            r = out.setdefault(label, dict(calls=0, total_ms=0.0, bytes=nbytes, flops=flops))
            r["calls"] += 1
            r["total_ms"] += ms
        for r in out.values():
            r["avg_us"] = 1e3 * r["total_ms"] / r["calls"]
            sec = r["avg_us"] * 1e-6
            r["gbs"] = r["bytes"] / sec / 1e9 if sec > 0 else 0.0
            r["tflops"] = r["flops"] / sec / 1e12 if sec > 0 else 0.0
        return out


@contextlib.contextmanager
def region(label, nbytes, flops=0):
    t = _active
    if t is None:
        yield
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        t.records.append((label, int(nbytes), int(flops), e0, e1))

"""Per-call CUDA-event timing of the C-ABI entry points (used by bench.py for the roofline rows).

When a ``KernelTimer`` is active, every call into libofr_b200.so is bracketed by two CUDA events
recorded on the stream the kernels are launched on (torch's current stream), tagged with the
call's algorithmic bytes and FLOPs (SURVEY.md section 8d formulas).  Nothing is synchronised until
``summary()``.  With no active timer the cost is one global lookup per call.
"""
import contextlib
import os
import statistics
import subprocess
import tempfile
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while a timed region runs (B200_PROFILING.md's
    clocks line); ``stop()`` returns {"sm_mhz": median, "sm_max_mhz", "reasons": [...], "samples"}."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out

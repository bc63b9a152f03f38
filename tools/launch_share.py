#!/usr/bin/env python
"""Share of the step per kernel from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_share.py launches.csv [--marker loss_bwd_kernel] [--step -2]

The list covers several training steps; `--marker` names a kernel launched exactly once per step (the last hot-path
kernel of the backward pass), which splits the list into steps.  The chosen step (default: the one before last, well
after warm-up) is summarised: ofr_b200 kernels individually, everything else (cuDNN / ATen: the host model) grouped.
ncu serialises launches and flushes caches, so only the SHARES are comparable with the CUDA-event timings of bench.py.
"""
import argparse
import collections
import csv
import re

OURS = ("corr_", "resample_", "chnorm_", "loss_", "polyphase_")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--marker", default="loss_bwd_kernel")
    ap.add_argument("--step", type=int, default=-2)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    launches = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            launches.append((r[ki], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
    steps, cur = [], []
    for name, ns in launches:
        cur.append((name, ns))
        if args.marker in name:
            steps.append(cur)
            cur = []
    print(f"{len(launches)} launches, {len(steps)} complete steps (marker {args.marker})")
    st = steps[args.step]
    tot = sum(ns for _, ns in st)
    ours = collections.OrderedDict()
    other = 0.0
    for name, ns in st:
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
        if any(t in short for t in OURS):
            key = re.sub(r"<.*", "", short)
            ours.setdefault(key, [0, 0.0])
            ours[key][0] += 1
            ours[key][1] += ns
        else:
            other += ns
    print(f"step {args.step}: {len(st)} launches, {tot / 1e6:.2f} ms of kernel time under ncu")
    print(f"  host model (cuDNN / ATen)          : {other / 1e6:8.2f} ms  {100 * other / tot:5.1f}%")
    osum = sum(v[1] for v in ours.values())
    print(f"  ofr_b200 kernels                   : {osum / 1e6:8.2f} ms  {100 * osum / tot:5.1f}%")
    for k, (n, ns) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"    {k:34s} x{n:3d} {ns / 1e3:9.1f} us  {100 * ns / tot:5.2f}% of step  {100 * ns / osum:5.1f}% of hot path")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Stall-sample shares per SASS address range of one kernel of an .ncu-rep.

    python tools/ncu_regions.py REPORT LAUNCH_INDEX [split offsets in hex ...]
"""
import collections
import csv
import io
import subprocess
import sys

rep, launch = sys.argv[1], sys.argv[2]
splits = sorted(int(x, 16) for x in sys.argv[3:])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", launch,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    data.append(r)
base = int(data[0][0], 16)
stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
reg = collections.defaultdict(collections.Counter)
ns = 0.0


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for r in data:
    off = int(r[0], 16) - base
    k = sum(1 for s in splits if off >= s)
    s = f(r[col["# Samples"]])
    ns += s
    for n in stall:
        reg[k][n] += f(r[col[n]])
    reg[k]["#"] += s
    reg[k]["inst"] += f(r[col["Instructions Executed"]])
    reg[k]["rows"] += 1
print("kernel:", rows[0][1][:100] if len(rows[0]) > 1 else "")
print("total samples", ns, "last offset", hex(int(data[-1][0], 16) - base))
bounds = [0] + splits + [int(data[-1][0], 16) - base + 16]
for k in sorted(reg):
    c = reg[k]
    print(f"[{bounds[k]:#07x},{bounds[k+1]:#07x}) rows {c['rows']:.0f} samples {c['#']:.0f} ({100 * c['#'] / ns:.1f}%) "
          f"warp-inst {c['inst']:.3g}")
    print("      " + ", ".join(f"{n[6:]} {v:.0f}" for n, v in sorted(c.items(), key=lambda x: -x[1])
                               if n.startswith("stall_") and v > 0.02 * c["#"]))

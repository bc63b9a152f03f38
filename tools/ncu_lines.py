#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions of one kernel launch in an .ncu-rep (needs -lineinfo).

    python tools/ncu_lines.py REPORT LAUNCH_INDEX [TOP_N]
"""
import collections
import csv
import io
import subprocess
import sys

rep, launch = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line #")]
print(rows[0][:3] if rows else "")
h = rows[hi[0]]
col = {n: i for i, n in enumerate(h)}
print(h[:8])

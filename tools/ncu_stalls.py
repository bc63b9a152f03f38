#!/usr/bin/env python
"""Warp-stall samples of the kernel launches in an .ncu-rep (`ncu --set full --import-source on`): the share of every stall
reason over all SASS instructions, and the instructions that collected the most samples.

    python tools/ncu_stalls.py REPORT [TOP_N]  >> profiles/rNN_ncu_<what>.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the page repeats per launch: a "Kernel Name" row, a header row, then one row per instruction
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)

seen = set()
for b in blocks:
    key = (b["name"], tuple(tuple(r[:6]) for r in b["rows"][:50]))
    if key in seen:   # the CSV page lists every launch once per source view
        continue
    seen.add(key)
    ix = {h: i for i, h in enumerate(b["hdr"])}

    def val(r, k):
        try:
            return float(r[ix[k]])
        except (KeyError, ValueError, IndexError):
            return 0.0

    total = sum(val(r, "# Samples") for r in b["rows"])
    print(f"## {b['name'][:100]}: {int(total)} warp-stall samples over {len(b['rows'])} SASS instructions")
    if total == 0:
        continue
    reasons = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    shares = sorted(((sum(val(r, h) for r in b["rows"]), h) for h in reasons), reverse=True)
    print("   " + ", ".join(f"{h[6:]} {100 * s / total:.1f} %" for s, h in shares if s / total >= 0.01))
    for r in sorted(b["rows"], key=lambda r: -val(r, "# Samples"))[:top]:
        worst = max(reasons, key=lambda h: val(r, h))
        print(f"   {100 * val(r, '# Samples') / total:5.1f} %  {r[ix['Source']].strip()[:80]:80s}  mostly {worst[6:]}")

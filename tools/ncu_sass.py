#!/usr/bin/env python
"""Per-kernel SASS digest of an .ncu-rep: executed warp instructions by opcode and the hottest stall sites.

    python tools/ncu_sass.py REPORT.ncu-rep LAUNCH_INDEX [--top N] [--view cuda|sass]
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("launch", type=int)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--view", default="sass")
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", args.view,
                          "--launch-skip", str(args.launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # locate the header row (starts with "Address" for sass, "Line No"-like for cuda)
    hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line No") and len(r) > 5)
    print(rows[0][:2])
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]

    def num(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return 0.0

    tot_inst = sum(num(r, "Instructions Executed") for r in data)
    tot_samp = sum(num(r, "# Samples") for r in data)
    print(f"rows {len(data)}  warp instructions executed {tot_inst:.3e}  stall samples {tot_samp:.0f}")
    if args.view == "sass":
        ops = collections.Counter()
        for r in data:
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
            op = m.group(1) if m else "?"
            op = ".".join(op.split(".")[:2])
            ops[op] += num(r, "Instructions Executed")
        print("-- opcode mix (share of executed warp instructions)")
        for op, n in ops.most_common(22):
            print(f"   {op:22s} {100 * n / max(tot_inst, 1):6.2f}%")
    print("-- hottest sites by stall samples")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:args.top]:
        s = sorted(((num(r, h), h[6:]) for h in stall_cols), reverse=True)[:3]
        why = " ".join(f"{h}:{int(v)}" for v, h in s if v > 0)
        src = r[col["Source"]].strip()[:90]
        key = r[0][-5:] if args.view == "sass" else r[0]
        print(f"   {100 * num(r, '# Samples') / max(tot_samp, 1):5.1f}%  x{num(r, 'Instructions Executed'):.2e}  [{key}] {src}   | {why}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""One-paragraph summary per kernel launch of an .ncu-rep (`ncu --set full`), the numbers DESIGN.md quotes.

    python tools/ncu_summary.py REPORT [REPORT ...] > profiles/rNN_ncu_<what>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [("dur us", "gpu__time_duration.sum"), ("dram R MB", "dram__bytes_read.sum"), ("dram W MB", "dram__bytes_write.sum"),
        ("regs", "launch__registers_per_thread"), ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("smem conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("warp-instr", "smsp__inst_executed.sum"), ("cycles", "sm__cycles_elapsed.max")]

for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")[:90]
        print(f"## {name}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        parts = []
        for label, key in KEYS:
            v = d.get(key)
            if v in (None, ""):
                continue
            try:
                f = float(v.replace(",", ""))
                unit = u.get(key, "")
                if label.endswith("MB"):  # ncu picks the unit per value
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}.get(unit, 1.0)
                if label == "dur us":
                    f *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(unit, 1.0)
                v = f"{f:.4g}" if f < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            parts.append(f"{label} {v}")
        print("   " + " | ".join(parts))
        st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
              for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
        print("   stalls/issue: " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))

#!/usr/bin/env python
"""Summarise an .ncu-rep (via `ncu -i ... --page raw --csv`) into a compact per-kernel table with stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
def f(r, k):
    try: return float(r[col[k]].replace(',', ''))
    except Exception: return float('nan')
for r in data:
    name = r[col['Kernel Name']]
    name = name.replace('void ', '').replace('<unnamed>::', '')[:70]
    print(f"## {name}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
    print(f"   dur {f(r,'gpu__time_duration.sum'):.1f} {units[col['gpu__time_duration.sum']]} | dram R {f(r,'dram__bytes_read.sum'):.1f} W {f(r,'dram__bytes_write.sum'):.1f} {units[col['dram__bytes_read.sum']]} | regs {r[col['launch__registers_per_thread']]} | "
          f"occupancy {f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}% | issue {f(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% | fma pipe {f(r,'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f}% | "
          f"smem wavefronts {f(r,'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')/1e6:.2f}M (conflicts {f(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')/1e6:.2f}M) | cycles {f(r,'sm__cycles_elapsed.avg')/1e3:.1f}K | eligible/cyc {f(r,'smsp__warps_eligible.avg.per_cycle_active'):.2f}")
    vals = sorted(((f(r, h), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stall), reverse=True)
    print("   stalls/issue: " + ", ".join(f"{h} {v:.2f}" for v, h in vals[:6]))

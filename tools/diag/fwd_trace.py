"""Diagnostic: clock64() timeline of CTA 0 of the PWC correlation forward (a -DOFR_FWD_TRACE=1 build of the library,
tools/build_variant.py; run with OFR_B200_LIB=<variant>).  Columns per ring iteration (one 8-channel chunk), in cycles:
wait = time at the stage's full barrier, fma = the chunk's arithmetic, bar = CTA barrier, issue = re-arming the stage
(thread 0 only), epi = the tile epilogue (last chunk of a tile)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ofr_b200

B, C, H = 24, 32, 128
f1 = torch.randn(B, C, H, H, device="cuda")
f2 = torch.randn(B, C, H, H, device="cuda")
corr = ofr_b200.Correlation(4, 1, 4, 1, 1, 1)
for _ in range(3):
    out = corr(f1, f2)
torch.cuda.synchronize()
lib = ofr_b200._lib.lib()
buf = (ctypes.c_longlong * (2 * 64 * 6))()
rc = lib.ofr_debug_fwd_trace(buf)
assert rc == 0, rc
for t in range(2):
    print(f"thread {t * 32}: q  start  wait  fma  bar  issue  epi   (cycles; start relative to q=0)")
    base = buf[t * 64 * 6]
    for q in range(64):
        r = [buf[(t * 64 + q) * 6 + i] for i in range(6)]
        if r[0] == 0:
            break
        epi = r[5] - r[4] if r[5] > r[4] else 0
        print(f"  {q:2d} {r[0] - base:8d} {r[1] - r[0]:6d} {r[2] - r[1]:6d} {r[3] - r[2]:6d} {r[4] - r[3]:6d} {epi:6d}")

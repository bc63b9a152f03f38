"""Diagnostic: clock64() marks of thread 0 of loss_final_kernel (a -DOFR_LOSS_TRACE=1 build of the library,
tools/build_variant.py; run with OFR_B200_LIB=<variant>) on the fused registration loss of the bench workload."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ofr_b200

B, S = 24, 512
sizes = [S >> i for i in range(7)]
flows = [torch.randn(B, 2, s, s, device="cuda") for s in sizes]
moving, fixed = torch.rand(B, 1, S, S, device="cuda"), torch.rand(B, 1, S, S, device="cuda")
for _ in range(3):
    out = ofr_b200.registration_loss(flows, moving, fixed)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
assert ofr_b200._lib.lib().ofr_debug_loss_trace(buf) == 0
names = ["start", "CTA partials folded by the thread groups", "barrier", "groups added, state written, barrier",
         "per-scale terms (float64 sqrt / div)", "barrier", "totals written"]
for i in range(1, 7):
    print(f"{names[i]:44s} +{buf[i] - buf[i - 1]:6d} cycles   (at {buf[i] - buf[0]})")

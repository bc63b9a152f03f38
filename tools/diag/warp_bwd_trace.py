"""Diagnostic: clock64() marks of thread 0 of one CTA of resample_bwd_fused_kernel (PWC warp backward, level 2 shape;
a -DOFR_WARP_TRACE=1 build of the library, tools/build_variant.py; run with OFR_B200_LIB=<variant>)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ofr_b200

B, C, H = 24, 32, 128
torch.manual_seed(0)
x = torch.randn(B, C, H, H, device="cuda", requires_grad=True)
flow = (torch.randn(B, 2, H, H, device="cuda") * 1.5).requires_grad_(True)
for _ in range(3):
    out = ofr_b200.pwc_warp(x, flow)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
assert ofr_b200._lib.lib().ofr_debug_warp_trace(buf) == 0
names = ["start", "count pass over the source tiles", "scan", "fill pass", "sort", "entry list", "chunk 0 gather",
         "chunk 0 flow gradient", "chunk 1 gather", "chunk 1 flow gradient"]
for i in range(1, 10):
    if buf[i]:
        print(f"{names[i]:36s} +{buf[i] - buf[i - 1]:6d} cycles   (at {buf[i] - buf[0]})")

import sys, ctypes, torch
sys.path.insert(0, "/root/repo")
import ofr_b200
from ofr_b200 import _lib
sys.path.insert(0, "/root/repo/tools")
import opbench
L = _lib.lib()
for B in (24, 4, 2):
    cs = opbench.cases(B, 512, only="loss|stn")
    for name, nbytes, flops, fn in cs:
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"B={B:2d} {name:40s} {us:8.1f} us  {nbytes/us/1e3:8.1f} GB/s  ({nbytes/1e6:.1f} MB, same buffers every launch)")

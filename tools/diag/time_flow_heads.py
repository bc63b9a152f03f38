"""Diagnostic: the flow-head kernels against cuDNN on the decoder's two-output-channel layers at the bench
shapes (B=24, 512x512 input), forward and backward (input + weight + bias gradients)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn as nn
import ofr_b200

torch.backends.cudnn.benchmark = True
B, dev = 24, "cuda"
cases = []
for lvl, h, width in ((6, 8, 529), (5, 16, 661), (4, 32, 629), (3, 64, 597)):
    cases.append((f"deconv{lvl} 2->2 {h}^2", "d", 2, h))
    cases.append((f"upfeat{lvl} {width}->2 {h}^2", "d", width, h))
    cases.append((f"predict_flow{lvl} {width}->2 {h}^2", "c", width, h))
cases += [("predict_flow2 565->2 128^2", "c", 565, 128), ("dc_conv7 32->2 128^2", "c", 32, 128),
          ("deconv2 2->2 128^2", "d", 2, 128), ("deconv1 2->2 256^2", "d", 2, 256)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


tot = [0.0, 0.0, 0.0, 0.0]
for name, kind, C, h in cases:
    ref = (nn.Conv2d(C, 2, 3, 1, 1) if kind == "c" else nn.ConvTranspose2d(C, 2, 4, 2, 1)).to(dev)
    ours = (ofr_b200.FlowConv2d(C, 2, 3, 1, 1) if kind == "c" else ofr_b200.FlowUpsample2d(C, 2, 4, 2, 1)).to(dev)
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(B, C, h, h, device=dev, requires_grad=True)
    go = torch.randn_like(ref(x))
    row = []
    for mod in (ref, ours):
        tf = timed(lambda: mod(x))

        def fb():
            torch.autograd.grad(mod(x), [x, mod.weight, mod.bias], go)

        row += [tf, timed(fb) - tf]
    for i in range(4):
        tot[i] += row[i]
    print(f"{name:32s} cuDNN fwd {row[0]:7.1f} bwd {row[1]:7.1f} us | ofr_b200 fwd {row[2]:7.1f} bwd {row[3]:7.1f} us")
print(f"total: cuDNN fwd {tot[0]:.0f} + bwd {tot[1]:.0f} us, ofr_b200 fwd {tot[2]:.0f} + bwd {tot[3]:.0f} us")

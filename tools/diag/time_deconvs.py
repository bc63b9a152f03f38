"""Diagnostic: what the ConvTranspose2d flow / feature upsamplers and the 2-channel output convolutions of
the PWC-Net decoder cost on cuDNN at the bench shapes (B=24, 512x512), forward + backward, CUDA-graph timed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn as nn

torch.backends.cudnn.benchmark = True
B = 24
dev = "cuda"
cases = []
for lvl, h, width in ((6, 8, 529), (5, 16, 661), (4, 32, 629), (3, 64, 597)):
    cases.append((f"deconv{lvl} 2->2 {h}^2", nn.ConvTranspose2d(2, 2, 4, 2, 1), (B, 2, h, h)))
    cases.append((f"upfeat{lvl} {width}->2 {h}^2", nn.ConvTranspose2d(width, 2, 4, 2, 1), (B, width, h, h)))
    cases.append((f"predict_flow{lvl} {width}->2 {h}^2", nn.Conv2d(width, 2, 3, 1, 1), (B, width, h, h)))
cases.append(("predict_flow2 565->2 128^2", nn.Conv2d(565, 2, 3, 1, 1), (B, 565, 128, 128)))
cases.append(("dc_conv7 32->2 128^2", nn.Conv2d(32, 2, 3, 1, 1), (B, 32, 128, 128)))
cases.append(("deconv2 2->2 128^2", nn.ConvTranspose2d(2, 2, 4, 2, 1), (B, 2, 128, 128)))
cases.append(("deconv1 2->2 256^2", nn.ConvTranspose2d(2, 2, 4, 2, 1), (B, 2, 256, 256)))


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


total_f = total_b = 0.0
for name, mod, shape in cases:
    mod = mod.to(dev)
    x = torch.randn(shape, device=dev, requires_grad=True)
    y = mod(x)
    go = torch.randn_like(y)
    tf = timed(lambda: mod(x))

    def fb():
        out = mod(x)
        torch.autograd.grad(out, [x, mod.weight, mod.bias], go)

    tb = timed(fb) - tf
    total_f += tf
    total_b += tb
    mb = 4 * (x.numel() + y.numel()) / 1e6
    print(f"{name:32s} fwd {tf:8.1f} us   bwd {tb:8.1f} us   in+out {mb:7.1f} MB")
print(f"total fwd {total_f:.0f} us, bwd {total_b:.0f} us")

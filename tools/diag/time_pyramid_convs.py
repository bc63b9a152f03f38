"""Diagnostic: what the thin first levels of the feature pyramid cost on cuDNN at the bench shapes (B=24 per image,
512x512), forward + backward, against the bytes they move."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn as nn

torch.backends.cudnn.benchmark = True
B, dev = 24, "cuda"
cases = [("conv1a 1->16 s2 512^2", 1, 16, 2, 512), ("conv1aa 16->16 256^2", 16, 16, 1, 256), ("conv1b 16->16 256^2", 16, 16, 1, 256),
         ("conv2a 16->32 s2 256^2", 16, 32, 2, 256), ("conv2aa 32->32 128^2", 32, 32, 1, 128), ("conv2b 32->32 128^2", 32, 32, 1, 128),
         ("conv3a 32->64 s2 128^2", 32, 64, 2, 128), ("conv3aa 64->64 64^2", 64, 64, 1, 64),
         ("conv2_4 533->32 128^2", 533, 32, 1, 128), ("conv2_3 469->64 128^2", 469, 64, 1, 128)]


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


tf_ = tb_ = 0.0
for name, cin, cout, stride, h in cases:
    m = nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1), nn.LeakyReLU(0.1)).to(dev)
    x = torch.randn(B, cin, h, h, device=dev, requires_grad=True)
    y = m(x)
    go = torch.randn_like(y)
    tf = timed(lambda: m(x))

    def fb():
        torch.autograd.grad(m(x), [x] + list(m.parameters()), go)

    tb = timed(fb) - tf
    mb = 4 * (x.numel() + y.numel()) / 1e6
    print(f"{name:28s} fwd {tf:7.1f} us  bwd {tb:7.1f} us   in+out {mb:7.1f} MB  ({mb / 6.5e3 * 1e3:.0f} us at the HBM peak)")
    tf_ += tf
    tb_ += tb
print(f"total fwd {tf_:.0f} bwd {tb_:.0f} us (each pyramid layer runs twice per step: two images)")

#!/usr/bin/env python
"""Top CUDA kernels of the bench's training step (torch.profiler), per host-model memory format.

    python tools/host_profile.py [--channels-last 0|1] [--steps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels-last", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--size", type=int, default=512)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    from ofr_b200.harness import make_optimizer, synthetic_pairs, train_step
    step, _ = bench.build_gpu_step(dev, channels_last=bool(args.channels_last))
    opt = make_optimizer(step.predictor)
    pair = synthetic_pairs(args.batch, args.size, seed=6).to(dev)
    for _ in range(4):
        train_step(step, opt, pair)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            train_step(step, opt, pair)
        torch.cuda.synchronize()
    rows = prof.key_averages()
    tot = sum(r.device_time_total for r in rows)
    print(f"channels_last={args.channels_last}: CUDA time {tot / args.steps / 1e3:.2f} ms/step")
    for r in sorted(rows, key=lambda r: -r.device_time_total)[:28]:
        print(f"  {100 * r.device_time_total / tot:5.1f}%  {r.device_time_total / args.steps / 1e3:8.3f} ms/step  x{r.count // args.steps:4d}  {r.key[:110]}")


if __name__ == "__main__":
    main()

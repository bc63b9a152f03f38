"""Harness around the hot path: the self-supervised registration training step and its synthetic data.

``RegistrationStep`` is the shape of the reference's hot loop (``train.py:41-57`` through
``opticalFlowReg.forward`` ``models.py:270-277`` and ``OFEloss`` ``loss.py:66-84``), minus the
segmentation/grid warps and host round-trips that only feed logging (``models.py:278-286``):

    flows  = predictor(pair)                         # host model, cuDNN convs + Correlation / pwc_warp
    warped = [stn(flow, moving) for flow in flows]   # one stn warp per scale
    loss   = OFEloss(flows, warped, fixed)[3]
    loss.backward(); optimizer.step()

The ``warp`` and ``loss`` callables are injected so the same step can be driven with the CUDA
layers (product) or, from ``bench.py`` / tests only, with the CPU oracle.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _smooth_field(shape, sigma, gen, device):
    """Gaussian-smoothed uniform noise (separable blur), shape [B,C,H,W]."""
    x = torch.rand(shape, generator=gen, device=device)
    radius = max(1, int(3 * sigma))
    t = torch.arange(-radius, radius + 1, device=device, dtype=torch.float32)
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = (k / k.sum()).view(1, 1, 1, -1)
    C = shape[1]
    x = F.conv2d(F.pad(x, (radius, radius, 0, 0), mode="reflect"), k.expand(C, 1, 1, -1), groups=C)
    x = F.conv2d(F.pad(x, (0, 0, radius, radius), mode="reflect"), k.transpose(2, 3).expand(C, 1, -1, 1), groups=C)
    return x


def synthetic_pairs(batch, size, seed=6, device="cpu", max_shift=8.0):
    """Synthetic fixed/moving slice pairs [B,2,S,S] in [0,1]: fixed = smooth random field
    min-max scaled per image (mirrors ScaleIntensityd, dataset.py:83); moving = fixed deformed by a smooth
    random flow of at most ``max_shift`` px (mirrors Rand2DElasticd, dataset.py:78).  Seed 6 = train.py:122."""
    gen = torch.Generator(device=device).manual_seed(seed)
    fixed = _smooth_field((batch, 1, size, size), 4.0, gen, device)
    lo = fixed.amin(dim=(2, 3), keepdim=True)
    hi = fixed.amax(dim=(2, 3), keepdim=True)
    fixed = (fixed - lo) / (hi - lo + 1e-12)
    disp = _smooth_field((batch, 2, size, size), 16.0, gen, device)
    disp = (disp - disp.mean(dim=(2, 3), keepdim=True))
    disp = disp / (disp.abs().amax(dim=(1, 2, 3), keepdim=True) + 1e-12) * max_shift
    ys, xs = torch.meshgrid(torch.arange(size, device=device, dtype=torch.float32),
                            torch.arange(size, device=device, dtype=torch.float32), indexing="ij")
    gx = 2 * (xs + disp[:, 0]) / (size - 1) - 1
    gy = 2 * (ys + disp[:, 1]) / (size - 1) - 1
    moving = F.grid_sample(fixed, torch.stack((gx, gy), -1), mode="bilinear", padding_mode="border",
                           align_corners=True)
    return torch.cat((fixed, moving), 1).contiguous()


class RegistrationStep(nn.Module):
    """predictor + per-scale stn warp + multi-scale loss; ``forward(pair)`` returns (p, c, s, total)."""

    def __init__(self, predictor, stn, loss_fn, stn_multi=None, reg_loss=None):
        """``stn_multi(flows, moving)`` (optional) replaces the per-scale ``stn`` loop with one batched call;
        ``reg_loss(flows, moving, fixed) -> (terms, warped)`` (optional) replaces warps + loss with the fused call."""
        super().__init__()
        self._reg_loss = reg_loss
        self.predictor = predictor
        self._stn = stn
        self._stn_multi = stn_multi
        self._loss = loss_fn

    def forward(self, pair):
        flows = list(self.predictor(pair))
        fixed = pair[:, :1].contiguous()
        moving = pair[:, 1:].contiguous()
        if self._reg_loss is not None:
            return self._reg_loss(flows, moving, fixed)[0]
        if self._stn_multi is not None:
            warped = self._stn_multi(flows, moving)
        else:
            warped = [self._stn(flow, moving) for flow in flows]
        return self._loss(flows, warped, fixed)


def make_optimizer(model, lr=1e-4, eps=1e-4, capturable=False):
    """Adam exactly as train.py:129 builds it (eps = lrMin default 1e-4).  ``capturable``: the step counter lives on
    the device so that the optimizer can be part of a captured CUDA graph (same arithmetic)."""
    params = [p for p in model.parameters() if p.requires_grad]
    return torch.optim.Adam(params, lr=lr, eps=eps, capturable=capturable)


def train_step(step, optimizer, pair):
    optimizer.zero_grad(set_to_none=True)
    p, c, s, total = step(pair)
    total.backward()
    optimizer.step()
    return total.detach()


class GraphedStep:
    """forward + loss + backward + Adam of one training step captured as ONE CUDA graph: possible
    because nothing on the hot path synchronises the host or allocates behind the caller's back (the reference's step
    cannot be captured: models.py:286 and loss.py:59 go through the host).  The batch is copied into a static device
    buffer; ``replay()`` returns the (static) loss tensor.  The optimizer must be built with ``capturable=True``."""

    def __init__(self, step, opt, example, warm_iters=3):
        self.static_in = example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warm_iters):  # (with DDP: at least 11 eager iterations before capture)
                train_step(step, opt, self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            total = step(self.static_in)[3]
            total.backward()
            opt.step()
            self.loss = total.detach()
        torch.cuda.synchronize()

    def replay(self, batch=None):
        if batch is not None:
            self.static_in.copy_(batch, non_blocking=True)
        self.graph.replay()
        return self.loss

"""Synthetic-deformation pair generator on the GPU.

The reference makes its training pairs on the host, one slice at a time, with MONAI (``dataset.py:72-85``):
``Resized(256)`` -> ``CopyItemsd`` -> ``Rand2DElasticd(prob=1, spacing=(16,16), magnitude_range=(0,0.5),
padding_mode="zeros", mode=[3,"nearest"])`` on the copy -> ``ConcatItemsd`` -> ``ScaleIntensityd(0,1)`` -- behind a
``num_workers=0`` loader (``dataset.py:31``), which becomes the bottleneck once the step itself is fast.

``ElasticPairGenerator`` does the same per BATCH on the device, reusing the Resample2d kernel for the warp:

1. control grid: MONAI's ``create_control_grid`` size rule (spacing 16 plus a border), offsets ``N(0,1) * m`` px with
   ``m ~ U(magnitude_range)`` per sample (``RandDeformGrid.randomize``);
2. dense displacement: bicubic upsampling by the spacing (``align_corners=False``), centre crop to the slice
   (``Rand2DElastic.__call__``);
3. moving = fixed sampled at ``p + d(p)``: ``ofr_resample_fwd`` (bilinear, border-clamped -- the reference uses a cubic
   spline with zero padding for the image; labels use the nearest mode, as the reference does);
4. joint min-max scaling of the (fixed, moving) pair to [0, 1] (``ScaleIntensityd`` on ``image_c``).

Everything is a device op on the current stream with a device ``torch.Generator``: no host round trip, reproducible
for a seed, and safe to run ahead of the training step on a side stream.
"""
import math

import torch
import torch.nn.functional as F

from .layers import Resample2d


def control_grid_size(size, spacing):
    """Points per axis of MONAI's control grid for an axis of ``size`` pixels (create_control_grid): enough points at
    ``spacing`` to cover the axis, symmetric about its centre, plus one border point on each side."""
    if size % 2 == 0:
        return int(math.ceil((size - 1.0) / (2.0 * spacing) + 0.5) * 2.0 + 2.0)
    return int(math.ceil((size - 1.0) / (2.0 * spacing)) * 2.0 + 3.0)


class ElasticPairGenerator:
    def __init__(self, size=256, spacing=16, magnitude_range=(0.0, 0.5), device="cuda", seed=6):
        self.size, self.spacing = int(size), int(spacing)
        self.magnitude_range = (float(magnitude_range[0]), float(magnitude_range[1]))
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self._warp = Resample2d()
        self._warp_nearest = Resample2d(bilinear=False)

    def displacement(self, batch):
        """[B,2,S,S] displacement field in pixels (x, y)."""
        n = control_grid_size(self.size, self.spacing)
        off = torch.randn(batch, 2, n, n, generator=self.gen, device=self.device)
        lo, hi = self.magnitude_range
        mag = lo + (hi - lo) * torch.rand(batch, 1, 1, 1, generator=self.gen, device=self.device)
        dense = F.interpolate(off * mag, scale_factor=self.spacing, mode="bicubic", align_corners=False)
        o = (dense.shape[-1] - self.size) // 2
        return dense[:, :, o:o + self.size, o:o + self.size].contiguous()

    def deform(self, fixed, segs=None):
        """fixed [B,1,S,S] (device, float32) -> pair [B,2,S,S] in [0,1]; with ``segs`` [B,1,S,S] also returns the label
        pair [B,2,S,S] (moving labels warped with the nearest mode)."""
        B = fixed.shape[0]
        flow = self.displacement(B)
        moving = self._warp(fixed, flow)
        pair = torch.cat((fixed, moving), 1)
        lo = pair.amin(dim=(1, 2, 3), keepdim=True)
        hi = pair.amax(dim=(1, 2, 3), keepdim=True)
        pair = (pair - lo) / (hi - lo).clamp_min(1e-12)
        if segs is None:
            return pair
        return pair, torch.cat((segs, self._warp_nearest(segs.to(torch.float32), flow)), 1)

    def __call__(self, batch):
        """A fresh synthetic batch: smooth random fixed images (the stand-in for the MRI slices, no dataset offline)
        and their elastic deformations."""
        from .harness.registration import _smooth_field
        fixed = _smooth_field((batch, 1, self.size, self.size), 4.0, self.gen, self.device)
        return self.deform(fixed)

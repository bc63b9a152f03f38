"""``opticalFlowReg`` (models.py:206-289) on the sm_100a kernels, without its host round trips.

The reference wrapper, per step: per-scale ``stn`` of the moving image (models.py:277), ``stn`` of the moving
segmentation and of a 256x256 line grid with the finest flow (``:278-281``), then
``torch.from_numpy(np.clip(np.rint(warped_segs.detach().cpu().numpy()), 0, 3)).to(device)`` (``:286``) -- a device->host
->device round trip and a synchronisation in every forward -- and ``grid_generator().view(...).to(device)`` (``:279``,
utils.py:15-23) rebuilt on the host every call.  Here:

* the line grid is built once per (size, device) and cached;
* segmentation and grid ride through ONE stn launch as a 2-channel frame (same flow);
* round-half-even + clamp run on the device (``torch.round`` == ``np.rint``), so nothing synchronises and the whole
  step can be captured in a CUDA graph.

``OpticalFlowReg(predictor)(x, segs)`` returns the reference's 4-tuple
``(flow_predictions, warped_images, warped_segs_int, warped_grid)``.
"""
import torch
import torch.nn as nn

from .layers import stn_warp, stn_warp_multi

_GRID_CACHE = {}


def grid_generator(size=256, device="cpu"):
    """utils.py:15-23: a ``size`` x ``size`` image of 1-px lines every 16 px starting at 7 (the reference hard-codes
    256; any size works here).  Cached per (size, device)."""
    key = (int(size), str(device))
    g = _GRID_CACHE.get(key)
    if g is None:
        g = torch.zeros(size, size)
        idx = torch.arange(7, size - 1, 16)
        g[idx, :] = 1.0
        g[:, idx] = 1.0
        g = g.view(1, 1, size, size).to(device)
        _GRID_CACHE[key] = g
    return g


def round_clamp_labels(warped_segs, lo=0, hi=3):
    """models.py:286 on the device: ``clip(rint(x), 0, 3)`` (round half to even, like numpy)."""
    return torch.round(warped_segs.detach()).clamp_(lo, hi)


class OpticalFlowReg(nn.Module):
    def __init__(self, predictor):
        super().__init__()
        self.predictor = predictor

    @staticmethod
    def stn(flow, frame):
        """models.py:256-268."""
        return stn_warp(flow, frame)

    def forward(self, x, segs=None):
        flows = list(self.predictor(x))
        moving = x[:, 1:2].contiguous()
        warped_images = stn_warp_multi(flows, moving)
        B, _, H, W = flows[0].shape
        grid = grid_generator(H, x.device) if H == W else None
        warped_segs_int = warped_grid = None
        with torch.no_grad():
            f0 = flows[0].detach()
            if segs is not None and grid is not None:
                frame = torch.cat((segs[:, 1:2].to(torch.float32), grid.expand(B, 1, H, W)), 1)
                both = stn_warp(f0, frame)
                warped_segs_int, warped_grid = round_clamp_labels(both[:, :1]), both[:, 1:]
            elif segs is not None:
                warped_segs_int = round_clamp_labels(stn_warp(f0, segs[:, 1:2].to(torch.float32).contiguous()))
            elif grid is not None:
                warped_grid = stn_warp(f0, grid.expand(B, 1, H, W).contiguous())
        return flows, warped_images, warped_segs_int, warped_grid

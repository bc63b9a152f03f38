"""Oracle (TEST INFRASTRUCTURE) for the two in-tree warps of the reference.

PINNED: ``tests/golden/warps.npz`` was produced by running the reference's own
``opticalFlowReg.stn`` (``models.py:256-268`` + ``generate_grid`` ``:195-204``)
and ``PWCDCNet.warp`` (``PWC/models/PWCNet.py:143-179``) unmodified; these
restatements are checked against it in ``tests/test_oracle.py``.

Both keep the reference's quirks on purpose (SURVEY.md section 7 "hard parts"):

* stn: the grid is scaled by 2/W (not 2/(W-1)) and sampled with
  ``align_corners=True``, so zero flow samples at x*(W-1)/W, not at x.
* pwc: the grid is scaled by 2/(W-1) and sampled with the default
  ``align_corners=False``, so zero flow samples at x*W/(W-1) - 0.5, and the
  0.9999-thresholded ones-mask zeroes every pixel with a tap outside the image.
"""
import torch
import torch.nn.functional as F


def pixel_grid(B, H, W, device, dtype):
    """[B,2,H,W] tensor of (x, y) pixel coordinates."""
    xs = torch.arange(W, device=device, dtype=dtype).view(1, 1, 1, W).expand(B, 1, H, W)
    ys = torch.arange(H, device=device, dtype=dtype).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat((xs, ys), dim=1)


def stn_warp(flow, frame, resize=True):
    """models.py:256-268.  ``frame`` [B,C,Hs,Ws] is first resized to the flow's
    (h,w) with bilinear/align_corners=True, then sampled at (p+flow)*(2/w,2/h)-1
    with ``grid_sample(bilinear, zeros, align_corners=True)``."""
    B, _, h, w = flow.shape
    if resize:
        frame = F.interpolate(frame, size=(h, w), mode="bilinear", align_corners=True)
    pts = (flow + pixel_grid(B, h, w, flow.device, flow.dtype)).permute(0, 2, 3, 1)
    # the reference builds the factor as a float32 tensor from python floats
    factor = torch.tensor([2.0 / w, 2.0 / h], dtype=torch.float32).to(flow.dtype).to(flow.device)
    grid = pts * factor.view(1, 1, 1, 2) - 1
    return F.grid_sample(frame, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def pwc_warp(x, flo):
    """PWC/models/PWCNet.py:143-179 (minus the device moves).  Returns ``output*mask``."""
    B, C, H, W = x.shape
    pts = pixel_grid(B, H, W, x.device, x.dtype) + flo
    gx = 2.0 * pts[:, 0] / max(W - 1, 1) - 1.0
    gy = 2.0 * pts[:, 1] / max(H - 1, 1) - 1.0
    grid = torch.stack((gx, gy), dim=-1)
    out = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    mask = F.grid_sample(torch.ones_like(x), grid, mode="bilinear", padding_mode="zeros",
                         align_corners=False)
    # thresholding is done in place on a fresh tensor in the reference: no gradient
    mask = (mask.detach() >= 0.9999).to(x.dtype)
    return out * mask

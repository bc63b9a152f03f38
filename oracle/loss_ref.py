"""Oracle (TEST INFRASTRUCTURE) for the self-supervised loss, ``loss.py:9-84``.

PINNED: ``tests/golden/loss.npz`` holds outputs of the reference's own
``photometric_loss`` / ``correlation_loss`` / ``smoothness_loss`` / ``OFEloss``
(imported unmodified, ``monai`` stubbed, ``tests/golden/make_golden.py``).

Kept quirks: the Charbonnier epsilon is squared (1e-9 -> 1e-18, loss.py:33-35);
the smoothness differences of the last row/column are taken against zero
(loss.py:25-26); NCC is a Pearson correlation over the WHOLE batch tensor,
divided once more by B, and short-circuits to 1 when either centred tensor is
identically zero (loss.py:55-62); scale weights 0.05*(i+1) live in float64 on
the host so every returned term is float64 (loss.py:71-73).
"""
import numpy as np
import torch
import torch.nn.functional as F


def charbonnier(x, alpha=0.25, epsilon=1.0e-9):
    return torch.pow(x * x + epsilon ** 2, alpha)


def _fixed_at(fixed, h, w):
    return F.interpolate(fixed, (h, w), mode="bilinear", align_corners=False)


def photometric(fixed, warped):
    """loss.py:9-13 -- sum of Charbonnier(fixed_resized - warped) / B."""
    h, w = warped.shape[2:]
    return charbonnier(_fixed_at(fixed, h, w) - warped).sum() / fixed.size(0)


def smoothness(flow):
    """loss.py:23-30 -- forward differences, last row/col against zero, /2 over the pair, /B."""
    B = flow.size(0)
    down = F.pad(flow[:, :, 1:, :], (0, 0, 0, 1))
    right = F.pad(flow[:, :, :, 1:], (0, 1, 0, 0))
    s = charbonnier(flow - down) + charbonnier(flow - right)
    return (s.sum(dim=1) / 2).sum() / B


def ncc(fixed, warped):
    """loss.py:52-64 -- 1 - (1/B) * Pearson(warped, fixed_resized) over the whole batch."""
    B, _, h, w = warped.shape
    f = _fixed_at(fixed, h, w)
    vx = warped - warped.mean()
    vy = f - f.mean()
    if not bool(vx.any()) or not bool(vy.any()):
        corr = torch.ones((), dtype=warped.dtype, device=warped.device)
    else:
        corr = (1.0 / B) * (vx * vy).sum() / (torch.sqrt((vx * vx).sum()) * torch.sqrt((vy * vy).sum()))
    return 1.0 - corr


def ofe_loss(flow, warped, fixed, lamb_da=0.5, gamma=100.0, zeta=100.0):
    """loss.py:66-84 -- returns (photometric, ncc, smoothness, total) as float64 0-dim tensors."""
    n = len(flow)
    weights = torch.from_numpy(0.05 * np.arange(1, n + 1))  # float64, stays on the host
    p = c = s = 0
    for i in range(n):
        p = p + weights[i] * photometric(fixed, warped[i])
        c = c + weights[i] * ncc(fixed, warped[i])
        s = s + weights[i] * smoothness(flow[i])
    p = 1 / n * gamma * p
    c = 1 / n * zeta * c
    s = 1 / n * lamb_da * s
    return p, c, s, p + s + c

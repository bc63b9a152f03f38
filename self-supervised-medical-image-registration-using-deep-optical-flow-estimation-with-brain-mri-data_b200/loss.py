"""``OFEloss`` and its terms on the fused sm_100a loss kernels (reference: ``loss.py:9-84``).

``OFEloss(flow, warped, fixed, lamb_da=0.5, gamma=100.0, zeta=100.0)`` keeps the reference's
signature and returns the same 4-tuple ``(p_loss, c_loss, s_loss, loss_all)`` of float64 0-dim
tensors (on the device: unlike the reference nothing is synchronised with the host).  Every
returned term is differentiable w.r.t. the ``warped`` images and the ``flow`` fields; ``fixed``
is treated as data (no gradient), which is how the reference uses it.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from .layers import _prep, _ptr, _stream, _workspace
from .profiling import region


def _pack_scales(flows, warpeds, g_warped=None, g_flow=None):
    n = len(flows)
    arr = (_lib.LossScale * n)()
    for i in range(n):
        h, w = warpeds[i].shape[2:]
        arr[i].warped = warpeds[i].data_ptr()
        arr[i].flow = flows[i].data_ptr()
        arr[i].g_warped = g_warped[i].data_ptr() if g_warped is not None and g_warped[i] is not None else None
        arr[i].g_flow = g_flow[i].data_ptr() if g_flow is not None and g_flow[i] is not None else None
        arr[i].h, arr[i].w = int(h), int(w)
    return arr


class _OFELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fixed, lamb_da, gamma, zeta, n, *tensors):
        flows = [_prep(t, f"OFEloss flow[{i}]") for i, t in enumerate(tensors[:n])]
        warpeds = [_prep(t, f"OFEloss warped[{i}]") for i, t in enumerate(tensors[n:])]
        fixed = _prep(fixed, "OFEloss fixed")
        B, cf, Hs, Ws = fixed.shape
        if cf != 1:
            raise ValueError("OFEloss: fixed must be [B,1,H,W]")
        for i in range(n):
            if warpeds[i].shape[0] != B or warpeds[i].shape[1] != 1:
                raise ValueError(f"OFEloss: warped[{i}] must be [B,1,h,w], got {tuple(warpeds[i].shape)}")
            if flows[i].shape != (B, 2) + tuple(warpeds[i].shape[2:]):
                raise ValueError(f"OFEloss: flow[{i}] {tuple(flows[i].shape)} does not match warped[{i}] "
                                 f"{tuple(warpeds[i].shape)}")
        dev = fixed.device
        out4 = torch.empty(4, dtype=torch.float64, device=dev)
        state = torch.empty(n * _lib.LOSS_STATE_DOUBLES, dtype=torch.float64, device=dev)
        px = sum(int(w.shape[2]) * int(w.shape[3]) for w in warpeds)
        with torch.cuda.device(dev), region(f"loss_fwd n{n} {Hs}x{Ws}", 4 * B * (3 * px + Hs * Ws)):
            ws_bytes = int(_lib.lib().ofr_loss_workspace(n))
            ws = _workspace(ws_bytes, dev)
            scales = _pack_scales(flows, warpeds)
            _lib.check(_lib.lib().ofr_loss_fwd(_ptr(fixed), B, Hs, Ws, scales, n, lamb_da, gamma, zeta,
                                               _ptr(out4), _ptr(state), _ptr(ws), ws_bytes, _stream()),
                       "ofr_loss_fwd")
        ctx.save_for_backward(fixed, state, *flows, *warpeds)
        ctx.cfg = (lamb_da, gamma, zeta, n)
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, gp, gc, gs, gt):
        lamb_da, gamma, zeta, n = ctx.cfg
        fixed, state = ctx.saved_tensors[:2]
        flows = list(ctx.saved_tensors[2:2 + n])
        warpeds = list(ctx.saved_tensors[2 + n:])
        B, _, Hs, Ws = fixed.shape
        dev = fixed.device
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        grad4 = torch.stack([(g if g is not None else zero).to(torch.float64).reshape(()) for g in (gp, gc, gs, gt)])
        need = ctx.needs_input_grad[5:]
        g_flow = [torch.empty_like(flows[i]) if need[i] else None for i in range(n)]
        g_warped = [torch.empty_like(warpeds[i]) if need[n + i] else None for i in range(n)]
        px = sum(int(w.shape[2]) * int(w.shape[3]) for w in warpeds)
        with torch.cuda.device(dev), region(f"loss_bwd n{n} {Hs}x{Ws}", 4 * B * (6 * px + Hs * Ws)):
            scales = _pack_scales(flows, warpeds, g_warped, g_flow)
            _lib.check(_lib.lib().ofr_loss_bwd(_ptr(fixed), B, Hs, Ws, scales, n, lamb_da, gamma, zeta,
                                               _ptr(grad4), _ptr(state), _stream()), "ofr_loss_bwd")
        return (None, None, None, None, None) + tuple(g_flow) + tuple(g_warped)


def OFEloss(flow, warped, fixed, lamb_da=0.5, gamma=100.0, zeta=100.0):
    """Drop-in for ``loss.OFEloss`` (loss.py:66-84)."""
    n = len(flow)
    if n != len(warped):
        raise ValueError("OFEloss: flow and warped must have the same number of scales")
    if not 1 <= n <= _lib.LOSS_MAX_SCALES:
        raise ValueError(f"OFEloss: between 1 and {_lib.LOSS_MAX_SCALES} scales are supported, got {n}")
    return _OFELossFn.apply(fixed, float(lamb_da), float(gamma), float(zeta), n, *flow, *warped)


class _RegLossFn(torch.autograd.Function):
    """stn of every scale + OFEloss in one pass each way (``ofr_regloss_fwd/bwd``)."""

    @staticmethod
    def forward(ctx, fixed, lamb_da, gamma, zeta, n, *tensors):
        movings = [_prep(t, f"registration_loss moving[{i}]") for i, t in enumerate(tensors[:n])]
        flows = [_prep(t, f"registration_loss flow[{i}]") for i, t in enumerate(tensors[n:])]
        fixed = _prep(fixed, "registration_loss fixed")
        B, cf, Hs, Ws = fixed.shape
        if cf != 1:
            raise ValueError("registration_loss: fixed must be [B,1,H,W]")
        for i in range(n):
            if flows[i].shape[0] != B or flows[i].shape[1] != 2:
                raise ValueError(f"registration_loss: flow[{i}] must be [B,2,h,w], got {tuple(flows[i].shape)}")
            if movings[i].shape != (B, 1) + tuple(flows[i].shape[2:]):
                raise ValueError(f"registration_loss: moving[{i}] {tuple(movings[i].shape)} does not match flow[{i}] "
                                 f"{tuple(flows[i].shape)}")
        dev = fixed.device
        out4 = torch.empty(4, dtype=torch.float64, device=dev)
        state = torch.empty(n * _lib.LOSS_STATE_DOUBLES, dtype=torch.float64, device=dev)
        warpeds = [torch.empty_like(m) for m in movings]
        px = sum(int(w.shape[2]) * int(w.shape[3]) for w in warpeds)
        # algorithmic bytes: moving + 2 flow read, warped written, fixed read once
        with torch.cuda.device(dev), region(f"regloss_fwd n{n} {Hs}x{Ws}", 4 * B * (4 * px + Hs * Ws)):
            ws_bytes = int(_lib.lib().ofr_loss_workspace(n))
            ws = _workspace(ws_bytes, dev)
            scales = _RegLossFn._scales(movings, flows, warpeds)
            _lib.check(_lib.lib().ofr_regloss_fwd(_ptr(fixed), B, Hs, Ws, scales, n, lamb_da, gamma, zeta,
                                                  _ptr(out4), _ptr(state), _ptr(ws), ws_bytes, _stream()),
                       "ofr_regloss_fwd")
        ctx.save_for_backward(fixed, state, *movings, *flows, *warpeds)
        ctx.cfg = (lamb_da, gamma, zeta, n)
        ctx.mark_non_differentiable(*warpeds)
        return (out4[0], out4[1], out4[2], out4[3]) + tuple(warpeds)

    @staticmethod
    def _scales(movings, flows, warpeds, g_flow=None):
        n = len(flows)
        arr = (_lib.RegLossScale * n)()
        for i in range(n):
            arr[i].moving = movings[i].data_ptr()
            arr[i].flow = flows[i].data_ptr()
            arr[i].warped = warpeds[i].data_ptr()
            arr[i].g_flow = g_flow[i].data_ptr() if g_flow is not None else None
            arr[i].h, arr[i].w = int(flows[i].shape[2]), int(flows[i].shape[3])
        return arr

    @staticmethod
    def backward(ctx, gp, gc, gs, gt, *_unused):
        lamb_da, gamma, zeta, n = ctx.cfg
        fixed, state = ctx.saved_tensors[:2]
        movings = list(ctx.saved_tensors[2:2 + n])
        flows = list(ctx.saved_tensors[2 + n:2 + 2 * n])
        warpeds = list(ctx.saved_tensors[2 + 2 * n:])
        B, _, Hs, Ws = fixed.shape
        dev = fixed.device
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        grad4 = torch.stack([(g if g is not None else zero).to(torch.float64).reshape(()) for g in (gp, gc, gs, gt)])
        g_flow = [torch.empty_like(f) for f in flows]
        px = sum(int(w.shape[2]) * int(w.shape[3]) for w in warpeds)
        # algorithmic bytes: warped + moving + 2 flow read, 2 flow-gradient planes written, fixed read once
        with torch.cuda.device(dev), region(f"regloss_bwd n{n} {Hs}x{Ws}", 4 * B * (6 * px + Hs * Ws)):
            scales = _RegLossFn._scales(movings, flows, warpeds, g_flow)
            _lib.check(_lib.lib().ofr_regloss_bwd(_ptr(fixed), B, Hs, Ws, scales, n, lamb_da, gamma, zeta,
                                                  _ptr(grad4), _ptr(state), _stream()), "ofr_regloss_bwd")
        need = ctx.needs_input_grad[5 + n:]
        return (None,) * (5 + n) + tuple(g if nd else None for g, nd in zip(g_flow, need))


def registration_loss(flows, moving, fixed, lamb_da=0.5, gamma=100.0, zeta=100.0):
    """``warped = [stn(flow, moving) for flow in flows]`` (models.py:277) and ``OFEloss(flows, warped, fixed)``
    (train.py:56, loss.py:66-84) fused: one launch forms every scale's warped image and its loss moments, one launch
    writes the complete flow gradients (smoothness + photometric/NCC chained through the warp).  Returns
    ``((p_loss, c_loss, s_loss, loss_all), warped_images)``; values and gradients are bit-identical to
    ``OFEloss(flows, stn_warp_multi(flows, moving), fixed)``.  The warped images come back detached (their
    gradient is already folded into the flows'); the reference only uses them for metrics and pictures afterwards.
    ``moving`` must not require a gradient (it is an input image)."""
    flows = list(flows)
    n = len(flows)
    if not 1 <= n <= _lib.LOSS_MAX_SCALES:
        raise ValueError(f"registration_loss: between 1 and {_lib.LOSS_MAX_SCALES} scales are supported, got {n}")
    if moving.requires_grad:
        raise ValueError("registration_loss: `moving` must not require a gradient; use stn_warp + OFEloss")
    movings = []
    for fl in flows:
        h, w = fl.shape[2:]
        movings.append(moving if moving.shape[2:] == (h, w) else
                       F.interpolate(moving, size=(h, w), mode="bilinear", align_corners=True))
    res = _RegLossFn.apply(fixed, float(lamb_da), float(gamma), float(zeta), n, *movings, *flows)
    return tuple(res[:4]), list(res[4:])


def _single_scale_terms(fixed, warped, flow):
    """(photometric, ncc, smoothness) of one scale as the reference's leaf functions return them: unweighted float32
    0-dim tensors (the kernels accumulate in float64; the cast is the only float32 rounding)."""
    # with n = 1 the weight is 0.05: choose coefficients so that the returned terms are unweighted
    p, c, s, _ = OFEloss([flow], [warped], fixed, lamb_da=20.0, gamma=20.0, zeta=20.0)
    return p.to(torch.float32), c.to(torch.float32), s.to(torch.float32)


def photometric_loss(fixed, warped):
    """loss.py:9-13."""
    flow = torch.zeros(warped.shape[0], 2, warped.shape[2], warped.shape[3], device=warped.device)
    return _single_scale_terms(fixed, warped, flow)[0]


def correlation_loss(fixed, warped):
    """loss.py:52-64."""
    flow = torch.zeros(warped.shape[0], 2, warped.shape[2], warped.shape[3], device=warped.device)
    return _single_scale_terms(fixed, warped, flow)[1]


def smoothness_loss(flow):
    """loss.py:23-30."""
    B, _, h, w = flow.shape
    img = torch.zeros(B, 1, h, w, device=flow.device)
    return _single_scale_terms(img, img, flow)[2]

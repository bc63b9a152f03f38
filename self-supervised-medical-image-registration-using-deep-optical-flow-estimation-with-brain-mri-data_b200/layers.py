"""The reference's custom-layer API, backed by the sm_100a kernels through the C ABI.

Module classes keep the flownet2 names and constructor signatures the reference model files use
(``flownet2/networks/FlowNetC.py:28,31``, ``flownet2/models.py:40,52``, ``PWC/models/PWCNet.py:69``):

    Correlation(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)(in1, in2)
    Resample2d(kernel_size=1, bilinear=True)(in1, flow)
    ChannelNorm(norm_deg=2)(in1)

plus the two in-tree warps as functions (``stn_warp`` = ``models.py:256-268``, ``pwc_warp`` =
``PWC/models/PWCNet.py:143-179``).  All are parameter-free, differentiable w.r.t. every tensor
input, float32 / CUDA only.  There is no CPU path: a CPU tensor raises.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .profiling import region
from ._lib import WARP_FLOWNET2, WARP_FLOWNET2_NEAREST, WARP_PWC, WARP_STN


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: ofr_b200 layers run on CUDA tensors only (got {t.device}); "
                           "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype} (wrap in tofp32()/tofp16() as the reference does)")
    if t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D NCHW tensor, got shape {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _workspace(nbytes, device):
    """256-byte aligned scratch from the caching allocator (the library itself never allocates)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------ Correlation
def correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    d, ho, wo = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _lib.lib().ofr_corr_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                       ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo))
    if rc != 0:
        raise ValueError(f"Correlation: invalid geometry H={H} W={W} pad={pad_size} k={kernel_size} "
                         f"md={max_displacement} s1={stride1} s2={stride2}")
    return d.value, ho.value, wo.value


class _CorrelationFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input1, input2, pad, k, md, s1, s2, slope=None):
        input1, input2 = _prep(input1, "Correlation input1"), _prep(input2, "Correlation input2")
        if input1.shape != input2.shape:
            raise ValueError(f"Correlation: input shapes differ {tuple(input1.shape)} vs {tuple(input2.shape)}")
        B, C, H, W = input1.shape
        D, Ho, Wo = correlation_out_shape(H, W, pad, k, md, s1, s2)
        out = torch.empty(B, D * D, Ho, Wo, dtype=torch.float32, device=input1.device)
        nbytes = 4 * (2 * B * C * H * W + B * D * D * Ho * Wo)
        with torch.cuda.device(input1.device), region(f"corr_fwd C{C} {H}x{W} D{D}", nbytes, 2 * B * D * D * C * Ho * Wo):
            ws_bytes = int(_lib.lib().ofr_corr_workspace(B, C, H, W, pad, k, md, s1, s2, 0))
            ws = _workspace(ws_bytes, input1.device) if ws_bytes else None
            _lib.check(_lib.lib().ofr_corr_fwd_act(_ptr(input1), _ptr(input2), _ptr(out), B, C, H, W, pad, k, md, s1, s2,
                                                   1.0 if slope is None else float(slope), _ptr(ws), ws_bytes,
                                                   _stream()), "ofr_corr_fwd")
        if slope is None:
            ctx.save_for_backward(input1, input2)
        else:  # the activation's backward needs the sign of the (activated) output
            ctx.save_for_backward(input1, input2, out)
        ctx.geom = (pad, k, md, s1, s2)
        ctx.slope = slope
        return out

    @staticmethod
    def backward(ctx, grad_out):
        input1, input2 = ctx.saved_tensors[:2]
        pad, k, md, s1, s2 = ctx.geom
        B, C, H, W = input1.shape
        if ctx.slope is not None:  # LeakyReLU backward: slope > 0, so sign(out) == sign(pre-activation)
            out = ctx.saved_tensors[2]
            grad_out = torch.where(out > 0, grad_out, grad_out * ctx.slope)
        grad_out = _prep(grad_out, "Correlation grad_output")
        g1 = torch.empty_like(input1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(input2) if ctx.needs_input_grad[1] else None
        if g1 is not None or g2 is not None:
            ng = (g1 is not None) + (g2 is not None)
            Dd = grad_out.shape[1]
            nbytes = 4 * (2 * B * C * H * W + grad_out.numel() + ng * B * C * H * W)
            with torch.cuda.device(input1.device), region(f"corr_bwd C{C} {H}x{W} D2={Dd} g{ng}", nbytes,
                                                          2 * ng * grad_out.numel() * C):
                ws_bytes = int(_lib.lib().ofr_corr_workspace(B, C, H, W, pad, k, md, s1, s2, 1))
                ws = _workspace(ws_bytes, input1.device) if ws_bytes else None
                _lib.check(_lib.lib().ofr_corr_bwd(_ptr(input1), _ptr(input2), _ptr(grad_out), _ptr(g1), _ptr(g2),
                                                   B, C, H, W, pad, k, md, s1, s2, _ptr(ws), ws_bytes, _stream()),
                           "ofr_corr_bwd")
        return g1, g2, None, None, None, None, None, None


class Correlation(nn.Module):
    """Cost volume.  Same constructor as flownet2's ``Correlation`` (defaults included);
    ``corr_multiply`` is accepted and ignored, as upstream."""

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1, *,
                 leaky_relu=None):
        """``leaky_relu`` (keyword only, not in flownet2's signature): negative slope of a LeakyReLU fused into the
        store of the cost volume -- the activation every reference call site applies next (PWCNet.py:201,
        FlowNetC.py:89).  ``None`` = the plain layer."""
        super().__init__()
        if leaky_relu is not None and not leaky_relu > 0:
            raise ValueError("Correlation: leaky_relu must be a positive slope")
        self.leaky_relu = leaky_relu
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return _CorrelationFn.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                    self.stride1, self.stride2, self.leaky_relu)

    def extra_repr(self):
        return (f"pad_size={self.pad_size}, kernel_size={self.kernel_size}, max_displacement={self.max_displacement}, "
                f"stride1={self.stride1}, stride2={self.stride2}")


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                               dilation_patch=1):
    """Alias of the ClementPinard sampler for the one configuration the reference calls
    (``FlowNetS/util.py:58-72``): un-normalised products, output ``(B, patch, patch, H, W)``."""
    if kernel_size != 1 or stride != 1 or padding != 0 or dilation != 1 or patch_size % 2 != 1:
        raise NotImplementedError("spatial_correlation_sample: only kernel_size=1, stride=1, padding=0, dilation=1 "
                                  "with an odd patch_size is provided (FlowNetS/util.py:58-66)")
    r = (patch_size - 1) // 2
    md = r * dilation_patch
    out = _CorrelationFn.apply(input1, input2, md, 1, md, 1, dilation_patch)
    B, _, H, W = out.shape
    return out.view(B, patch_size, patch_size, H, W) * float(input1.size(1))


# ------------------------------------------------------------------------------------------ warps
class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input1, flow, mode, flow_scale=1.0):
        input1, flow = _prep(input1, "warp input1"), _prep(flow, "warp flow")
        B, C, Hin, Win = input1.shape
        Bf, two, Hf, Wf = flow.shape
        if Bf != B or two != 2:
            raise ValueError(f"warp: flow must be [B,2,H,W] with B={B}, got {tuple(flow.shape)}")
        if mode in (WARP_STN, WARP_PWC) and (Hin, Win) != (Hf, Wf):
            raise ValueError("warp: image and flow must have the same spatial size")
        out = torch.empty(B, C, Hf, Wf, dtype=torch.float32, device=input1.device)
        with torch.cuda.device(input1.device), region(f"warp_fwd m{mode} C{C} {Hf}x{Wf}", 4 * B * Hf * Wf * (2 * C + 2)):
            _lib.check(_lib.lib().ofr_resample_fwd_scaled(_ptr(input1), _ptr(flow), _ptr(out), B, C, Hin, Win, Hf, Wf,
                                                          mode, float(flow_scale), _stream()), "ofr_resample_fwd")
        ctx.save_for_backward(input1, flow)
        ctx.mode = mode
        ctx.flow_scale = float(flow_scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        input1, flow = ctx.saved_tensors
        mode = ctx.mode
        B, C, Hin, Win = input1.shape
        _, _, Hf, Wf = flow.shape
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if mode == WARP_FLOWNET2_NEAREST:
            if need1:
                raise NotImplementedError("Resample2d(bilinear=False): image gradient is not provided")
            return None, (torch.zeros_like(flow) if need2 else None), None, None
        grad_out = _prep(grad_out, "warp grad_output")
        g1 = torch.empty_like(input1) if need1 else None
        g2 = torch.empty_like(flow) if need2 else None
        if g1 is None and g2 is None:
            return None, None, None, None
        nbytes = 4 * B * Hf * Wf * (2 * C + 2 + (C if g1 is not None else 0) + (2 if g2 is not None else 0))
        with torch.cuda.device(input1.device), region(
                f"warp_bwd m{mode} C{C} {Hf}x{Wf} gI1={int(g1 is not None)} gI2={int(g2 is not None)}", nbytes):
            ws, ws_bytes = None, 0
            if g1 is not None:
                ws_bytes = int(_lib.lib().ofr_resample_bwd_workspace(B, C, Hin, Win, Hf, Wf, mode))
                ws = _workspace(ws_bytes, input1.device)
            _lib.check(_lib.lib().ofr_resample_bwd_scaled(_ptr(input1), _ptr(flow), _ptr(grad_out), _ptr(g1), _ptr(g2),
                                                          B, C, Hin, Win, Hf, Wf, mode, ctx.flow_scale, _ptr(ws), ws_bytes,
                                                          _stream()), "ofr_resample_bwd")
        return g1, g2, None, None


class Resample2d(nn.Module):
    """Bilinear backward warp in pixel units, border-clamped taps (flownet2 ``Resample2d``)."""

    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError("Resample2d: only kernel_size=1 (the value every reference call site uses, "
                                      "flownet2/models.py:52-88) is provided")
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        return _WarpFn.apply(input1, input2, WARP_FLOWNET2 if self.bilinear else WARP_FLOWNET2_NEAREST)


def stn_warp(flow, frame):
    """``opticalFlowReg.stn(flow, frame)`` (models.py:256-268): resize ``frame`` to the flow's size
    (bilinear, align_corners=True), then sample it at (p+flow)*(2/w,2/h)-1 with zeros padding."""
    h, w = flow.shape[2:]
    if frame.shape[2:] != (h, w):
        frame = F.interpolate(frame, size=(h, w), mode="bilinear", align_corners=True)
    return _WarpFn.apply(frame, flow, WARP_STN)


class _StnMultiFn(torch.autograd.Function):
    """All per-scale stn warps in one launch each way; inputs (frame_0..frame_{n-1}, flow_0..flow_{n-1})."""

    @staticmethod
    def _scales(frames, flows, outs=None, gouts=None, gflows=None):
        n = len(flows)
        arr = (_lib.StnScale * n)()
        for i in range(n):
            arr[i].frame, arr[i].flow = frames[i].data_ptr(), flows[i].data_ptr()
            arr[i].out = outs[i].data_ptr() if outs is not None else None
            arr[i].g_out = gouts[i].data_ptr() if gouts is not None else None
            arr[i].g_flow = gflows[i].data_ptr() if gflows is not None else None
            arr[i].h, arr[i].w = flows[i].shape[2], flows[i].shape[3]
        return arr

    @staticmethod
    def forward(ctx, *tensors):
        n = len(tensors) // 2
        frames = [_prep(t, "stn frame") for t in tensors[:n]]
        flows = [_prep(t, "stn flow") for t in tensors[n:]]
        B, C = frames[0].shape[:2]
        for fr, fl in zip(frames, flows):
            if fr.shape[0] != B or fr.shape[1] != C or fl.shape[0] != B or fl.shape[1] != 2 or fr.shape[2:] != fl.shape[2:]:
                raise ValueError("stn_warp_multi: every frame must be [B,C,h,w] with a [B,2,h,w] flow")
        outs = [torch.empty_like(fr) for fr in frames]
        nbytes = sum(4 * B * fl.shape[2] * fl.shape[3] * (2 * C + 2) for fl in flows)
        with torch.cuda.device(frames[0].device), region(f"stn_multi_fwd n{n} C{C} {flows[0].shape[2]}x{flows[0].shape[3]}",
                                                         nbytes):
            arr = _StnMultiFn._scales(frames, flows, outs=outs)
            _lib.check(_lib.lib().ofr_stn_multi_fwd(arr, n, B, C, _stream()), "ofr_stn_multi_fwd")
        ctx.save_for_backward(*frames, *flows)
        ctx.n = n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        n = ctx.n
        frames, flows = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        B, C = frames[0].shape[:2]
        if not any(ctx.needs_input_grad[n:]):
            return (None,) * (2 * n)
        gouts = [_prep(g if g is not None else torch.zeros_like(fr), "stn grad_output") for g, fr in zip(gouts, frames)]
        gflows = [torch.empty_like(fl) for fl in flows]
        nbytes = sum(4 * B * fl.shape[2] * fl.shape[3] * (2 * C + 4) for fl in flows)
        with torch.cuda.device(frames[0].device), region(f"stn_multi_bwd n{n} C{C} {flows[0].shape[2]}x{flows[0].shape[3]}",
                                                         nbytes):
            arr = _StnMultiFn._scales(frames, flows, gouts=gouts, gflows=gflows)
            _lib.check(_lib.lib().ofr_stn_multi_bwd(arr, n, B, C, _stream()), "ofr_stn_multi_bwd")
        return (None,) * n + tuple(g if need else None for g, need in zip(gflows, ctx.needs_input_grad[n:]))


def stn_warp_multi(flows, frame):
    """``[stn_warp(flow, frame) for flow in flows]`` -- the per-scale loop of ``opticalFlowReg.forward``
    (models.py:277-281) -- with one kernel launch for all scales in each direction (most scales are a few thousand
    pixels: a launch per scale is pure latency).  The frame needs no gradient (it is an input image); a frame that
    requires one must use ``stn_warp``."""
    flows = list(flows)
    if len(flows) > _lib.STN_MAX_SCALES:
        raise ValueError(f"stn_warp_multi: at most {_lib.STN_MAX_SCALES} scales")
    if frame.requires_grad:
        return [stn_warp(fl, frame) for fl in flows]
    frames = []
    for fl in flows:
        h, w = fl.shape[2:]
        frames.append(frame if frame.shape[2:] == (h, w) else
                      F.interpolate(frame, size=(h, w), mode="bilinear", align_corners=True))
    return list(_StnMultiFn.apply(*frames, *flows))


def pwc_warp(x, flo, flow_scale=1.0):
    """``PWCDCNet.warp(x, flo)`` (PWC/models/PWCNet.py:143-179) in one pass: sample, build the
    ones-mask, threshold at 0.9999 and multiply, without materialising grid or mask.
    ``pwc_warp(x, flo, s)`` == ``pwc_warp(x, flo * s)`` (the call sites ``:214,228,242,258`` pre-scale the
    upsampled flow by 0.625 ... 5.0) without the extra elementwise pass."""
    return _WarpFn.apply(x, flo, WARP_PWC, flow_scale)


# ------------------------------------------------------------------------------------------ ChannelNorm
class _ChannelNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input1):
        input1 = _prep(input1, "ChannelNorm input1")
        B, C, H, W = input1.shape
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=input1.device)
        with torch.cuda.device(input1.device), region(f"chnorm_fwd C{C} {H}x{W}", 4 * B * H * W * (C + 1)):
            _lib.check(_lib.lib().ofr_chnorm_fwd(_ptr(input1), _ptr(out), B, C, H, W, _stream()), "ofr_chnorm_fwd")
        ctx.save_for_backward(input1, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        input1, out = ctx.saved_tensors
        B, C, H, W = input1.shape
        grad_out = _prep(grad_out, "ChannelNorm grad_output")
        g = torch.empty_like(input1)
        with torch.cuda.device(input1.device), region(f"chnorm_bwd C{C} {H}x{W}", 4 * B * H * W * (2 * C + 2)):
            _lib.check(_lib.lib().ofr_chnorm_bwd(_ptr(input1), _ptr(out), _ptr(grad_out), _ptr(g), B, C, H, W,
                                                 _stream()), "ofr_chnorm_bwd")
        return g


class _ChannelNormDiffFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _prep(a, "ChannelNorm input1"), _prep(b, "ChannelNorm input2")
        if a.shape != b.shape:
            raise ValueError(f"channelnorm_diff: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
        B, C, H, W = a.shape
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device), region(f"chnorm_diff_fwd C{C} {H}x{W}", 4 * B * H * W * (2 * C + 1)):
            _lib.check(_lib.lib().ofr_chnorm_diff_fwd(_ptr(a), _ptr(b), _ptr(out), B, C, H, W, _stream()),
                       "ofr_chnorm_diff_fwd")
        ctx.save_for_backward(a, b, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, b, out = ctx.saved_tensors
        B, C, H, W = a.shape
        grad_out = _prep(grad_out, "ChannelNorm grad_output")
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if ga is None and gb is None:
            return None, None
        n_out = (ga is not None) + (gb is not None)
        with torch.cuda.device(a.device), region(f"chnorm_diff_bwd C{C} {H}x{W}", 4 * B * H * W * ((2 + n_out) * C + 2)):
            _lib.check(_lib.lib().ofr_chnorm_diff_bwd(_ptr(a), _ptr(b), _ptr(out), _ptr(grad_out), _ptr(ga), _ptr(gb),
                                                      B, C, H, W, _stream()), "ofr_chnorm_diff_bwd")
        return ga, gb


def channelnorm_diff(a, b):
    """``ChannelNorm()(a - b)`` in one pass (the ``x[:, :3] - resampled`` -> ``channelnorm`` pairs of
    flownet2/models.py:140-141,152-153,166-167,179-180) without materialising the difference."""
    return _ChannelNormDiffFn.apply(a, b)


class ChannelNorm(nn.Module):
    """Per-pixel L2 norm over channels (flownet2 ``ChannelNorm``; ``norm_deg`` is ignored upstream too)."""

    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return _ChannelNormFn.apply(input1)

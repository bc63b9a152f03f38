"""Flow heads of the PWC-Net decoder on the sm_100a kernels of ``csrc/flowhead.cu``.

The reference builds its two-output-channel layers with two helpers (``PWC/models/PWCNet.py:30-34``)::

    predict_flow(in_planes)   = nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)
    deconv(in_planes, 2, ...) = nn.ConvTranspose2d(in_planes, 2, kernel_size=4, stride=2, padding=1, bias=True)

``FlowConv2d`` / ``FlowUpsample2d`` are subclasses of the two ``nn`` modules with the same constructor, parameters and
state dict; their ``forward`` goes through the C ABI (``ofr_flowconv3_*`` / ``ofr_flowdeconv4_*``) when the layer has
that shape and the input is a float32 CUDA tensor whose width is a multiple of 4, and through the parent class (the
reference's own cuDNN call) otherwise.  ``accelerate_flow_heads(model)`` re-classes the matching layers of an already
constructed model in place -- the reference's ``PWCDCNet`` needs no edit and keeps loading its checkpoints.
"""
import torch
import torch.nn as nn

from . import _lib
from .layers import _ptr, _stream, _workspace
from .profiling import region


def _fits(x, weight):
    """Inputs the kernels take; anything else (CPU tensors, other dtypes, widths that are not a multiple of 4) stays with the
    parent class."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[3] % 4 == 0 and
            weight.dtype == torch.float32 and weight.is_cuda and x.numel() > 0)


class _FlowConv3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = x.contiguous(), weight.contiguous()
        B, C, H, W = x.shape
        out = torch.empty(B, 2, H, W, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _workspace(L.ofr_flowconv3_workspace(B, C, H, W, 0), x.device)
        with torch.cuda.device(x.device), region(f"flowconv3_fwd C{C} {H}x{W}", 4 * B * H * W * (C + 2), 2 * 18 * B * C * H * W):
            _lib.check(L.ofr_flowconv3_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(out), B, C, H, W, _ptr(ws), ws.numel(),
                                           _stream()), "ofr_flowconv3_fwd")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        B, C, H, W = x.shape
        grad_out = grad_out.contiguous()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(weight) if need_w else None
        gb = torch.empty(2, dtype=torch.float32, device=x.device) if need_w and ctx.has_bias else None
        L = _lib.lib()
        ws = _workspace(L.ofr_flowconv3_workspace(B, C, H, W, 1), x.device)
        with torch.cuda.device(x.device), region(f"flowconv3_bwd C{C} {H}x{W}", 4 * B * H * W * (2 * C + 2), 4 * 18 * B * C * H * W):
            _lib.check(L.ofr_flowconv3_bwd(_ptr(x), _ptr(weight), _ptr(grad_out), _ptr(gx), _ptr(gw), _ptr(gb), B, C, H, W,
                                           _ptr(ws), ws.numel(), _stream()), "ofr_flowconv3_bwd")
        return gx, gw, gb


class _FlowDeconv4Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = x.contiguous(), weight.contiguous()
        B, C, H, W = x.shape
        out = torch.empty(B, 2, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _workspace(L.ofr_flowdeconv4_workspace(B, C, H, W, 0), x.device)
        with torch.cuda.device(x.device), region(f"flowdeconv4_fwd C{C} {H}x{W}", 4 * B * H * W * (C + 8), 2 * 32 * B * C * H * W):
            _lib.check(L.ofr_flowdeconv4_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(out), B, C, H, W, _ptr(ws), ws.numel(),
                                             _stream()), "ofr_flowdeconv4_fwd")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        B, C, H, W = x.shape
        grad_out = grad_out.contiguous()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(weight) if need_w else None
        gb = torch.empty(2, dtype=torch.float32, device=x.device) if need_w and ctx.has_bias else None
        L = _lib.lib()
        ws = _workspace(L.ofr_flowdeconv4_workspace(B, C, H, W, 1), x.device)
        with torch.cuda.device(x.device), region(f"flowdeconv4_bwd C{C} {H}x{W}", 4 * B * H * W * (2 * C + 8), 4 * 32 * B * C * H * W):
            _lib.check(L.ofr_flowdeconv4_bwd(_ptr(x), _ptr(weight), _ptr(grad_out), _ptr(gx), _ptr(gw), _ptr(gb), B, C, H, W,
                                             _ptr(ws), ws.numel(), _stream()), "ofr_flowdeconv4_bwd")
        return gx, gw, gb


def _is_flow_conv(m):
    return (type(m) in (nn.Conv2d, FlowConv2d) and m.out_channels == 2 and m.kernel_size == (3, 3) and m.stride == (1, 1) and
            m.padding == (1, 1) and m.dilation == (1, 1) and m.groups == 1 and m.padding_mode == "zeros")


def _is_flow_deconv(m):
    return (type(m) in (nn.ConvTranspose2d, FlowUpsample2d) and m.out_channels == 2 and m.kernel_size == (4, 4) and
            m.stride == (2, 2) and m.padding == (1, 1) and m.output_padding == (0, 0) and m.dilation == (1, 1) and
            m.groups == 1)


class FlowConv2d(nn.Conv2d):
    """``predict_flow`` (``PWC/models/PWCNet.py:30-31``): ``nn.Conv2d(C, 2, 3, 1, 1)`` with the 2-channel kernels."""

    def forward(self, x):
        if _is_flow_conv(self) and _fits(x, self.weight):
            return _FlowConv3Fn.apply(x, self.weight, self.bias)
        return super().forward(x)


class FlowUpsample2d(nn.ConvTranspose2d):
    """``deconv`` (``PWC/models/PWCNet.py:33-34``): ``nn.ConvTranspose2d(C, 2, 4, 2, 1)`` with the 2-channel kernels."""

    def forward(self, x, output_size=None):
        if output_size is None and _is_flow_deconv(self) and _fits(x, self.weight):
            return _FlowDeconv4Fn.apply(x, self.weight, self.bias)
        return super().forward(x, output_size)


def accelerate_flow_heads(model):
    """Re-class, in place, every ``nn.Conv2d(C, 2, 3, 1, 1)`` / ``nn.ConvTranspose2d(C, 2, 4, 2, 1)`` of ``model`` (the
    reference's ``predict_flow*``, ``dc_conv7``, ``deconv*``, ``upfeat*``) to ``FlowConv2d`` / ``FlowUpsample2d``.
    Parameters, buffers and state-dict keys are untouched.  Returns the names of the layers it changed."""
    changed = []
    for name, m in model.named_modules():
        if type(m) is nn.Conv2d and _is_flow_conv(m):
            m.__class__ = FlowConv2d
            changed.append(name)
        elif type(m) is nn.ConvTranspose2d and _is_flow_deconv(m):
            m.__class__ = FlowUpsample2d
            changed.append(name)
    return changed

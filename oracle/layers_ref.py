"""Oracle (TEST INFRASTRUCTURE) for the three un-vendored custom layers.

PARITY UNPINNED against NVIDIA/flownet2-pytorch: the reference tree only imports
these layers (``flownet2/networks/FlowNetC.py:8,28,31,88``;
``flownet2/models.py:10-11,40,52,139-180``; ``PWC/models/PWCNet.py:13,69,200-259``;
``FlowNetS/util.py:58-72``).  The definitions below are the ones SURVEY.md
section 8a fixes, cross-checked against the TVM numpy correlation and against
``F.grid_sample`` (see ``tests/test_oracle.py``).

Every function is dtype-generic: run it in float64 for a ground truth, in
float32 for a like-for-like comparison.  Gradients come from autograd of these
restatements, except ChannelNorm whose upstream backward carries an explicit
``+1e-8`` and is therefore written out.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    """(D, Hout, Wout) of the cost volume -- SURVEY.md 8a row a1."""
    kr = (kernel_size - 1) // 2
    border = max_displacement + kr
    Hout = int(math.ceil((H + 2 * pad_size - 2 * border) / float(stride1)))
    Wout = int(math.ceil((W + 2 * pad_size - 2 * border) / float(stride1)))
    D = 2 * (max_displacement // stride2) + 1
    return D, Hout, Wout


def correlation(f1, f2, pad_size, kernel_size, max_displacement, stride1=1, stride2=1):
    """Cost volume, shift-and-multiply form.

    out[b, (tj+r)*D + (ti+r), y, x] =
        1/(k*k*C) * sum_{c,h,w} P1[b,c,y1+h,x1+w] * P2[b,c,y1+tj*s2+h, x1+ti*s2+w]
    with P = zero padded input, y1 = y*s1 + md, h,w in [0,k)  (Caffe convention;
    the reference only ever uses k=1: FlowNetC.py:31, PWCNet.py:69).
    """
    B, C, H, W = f1.shape
    D, Hout, Wout = correlation_out_shape(H, W, pad_size, kernel_size, max_displacement,
                                          stride1, stride2)
    if Hout < 1 or Wout < 1:
        raise ValueError("correlation: empty output")
    r = max_displacement // stride2
    k = kernel_size
    p1 = F.pad(f1, (pad_size,) * 4)
    p2 = F.pad(f2, (pad_size,) * 4)
    # the windows read up to y1 + k - 1; give slack so ceil() rounding never over-reads
    need_h = (Hout - 1) * stride1 + max_displacement + r * stride2 + k
    need_w = (Wout - 1) * stride1 + max_displacement + r * stride2 + k
    eh = max(0, need_h - p1.shape[2])
    ew = max(0, need_w - p1.shape[3])
    if eh or ew:
        p1 = F.pad(p1, (0, ew, 0, eh))
        p2 = F.pad(p2, (0, ew, 0, eh))
    outs = []
    ys = max_displacement
    xs = max_displacement
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            acc = 0
            for h in range(k):
                for w in range(k):
                    a = p1[:, :, ys + h: ys + h + (Hout - 1) * stride1 + 1: stride1,
                           xs + w: xs + w + (Wout - 1) * stride1 + 1: stride1]
                    y2 = ys + tj * stride2 + h
                    x2 = xs + ti * stride2 + w
                    b = p2[:, :, y2: y2 + (Hout - 1) * stride1 + 1: stride1,
                           x2: x2 + (Wout - 1) * stride1 + 1: stride1]
                    acc = acc + (a * b).sum(dim=1)
            outs.append(acc)
    out = torch.stack(outs, dim=1)
    return out / float(k * k * C)


def correlation_loops(f1, f2, pad_size, kernel_size, max_displacement, stride1=1, stride2=1):
    """Scalar-loop numpy statement of the same operator (tiny inputs only)."""
    f1 = np.asarray(f1, dtype=np.float64)
    f2 = np.asarray(f2, dtype=np.float64)
    B, C, H, W = f1.shape
    D, Hout, Wout = correlation_out_shape(H, W, pad_size, kernel_size, max_displacement,
                                          stride1, stride2)
    r = max_displacement // stride2

    def at(f, b, c, y, x):  # padded-coordinate read, zero outside the image
        y -= pad_size
        x -= pad_size
        if 0 <= y < H and 0 <= x < W:
            return f[b, c, y, x]
        return 0.0

    out = np.zeros((B, D * D, Hout, Wout))
    for b in range(B):
        for y in range(Hout):
            for x in range(Wout):
                y1 = y * stride1 + max_displacement
                x1 = x * stride1 + max_displacement
                for tj in range(-r, r + 1):
                    for ti in range(-r, r + 1):
                        s = 0.0
                        for h in range(kernel_size):
                            for w in range(kernel_size):
                                for c in range(C):
                                    s += at(f1, b, c, y1 + h, x1 + w) * \
                                         at(f2, b, c, y1 + tj * stride2 + h, x1 + ti * stride2 + w)
                        out[b, (tj + r) * D + (ti + r), y, x] = s
    return out / float(kernel_size * kernel_size * C)


def resample2d(in1, flow, bilinear=True):
    """Backward warp in pixel units with border-clamped taps -- SURVEY.md 8a row a4.

    out[b,c,y,x] = sum over the 2x2 taps of in1[b,c] around (x+u, y+v); tap indices
    are clamped to the image, weights are the *unclamped* fractional parts.  The
    output takes the spatial size of ``flow`` (call sites flownet2/models.py:139,
    151,163,176 always pass equal sizes).
    """
    B, C, H, W = in1.shape
    _, _, Hf, Wf = flow.shape
    dev, dt = in1.device, in1.dtype
    xs = torch.arange(Wf, device=dev, dtype=dt).view(1, 1, Wf)
    ys = torch.arange(Hf, device=dev, dtype=dt).view(1, Hf, 1)
    xf = xs + flow[:, 0]
    yf = ys + flow[:, 1]
    flat = in1.reshape(B, C, H * W)

    def gather(yi, xi):
        idx = (yi * W + xi).view(B, 1, Hf * Wf).expand(B, C, Hf * Wf)
        return torch.gather(flat, 2, idx).view(B, C, Hf, Wf)

    if not bilinear:
        xn = torch.floor(xf + 0.5).long().clamp(0, W - 1)
        yn = torch.floor(yf + 0.5).long().clamp(0, H - 1)
        return gather(yn, xn)
    x0 = torch.floor(xf)
    y0 = torch.floor(yf)
    a = (xf - x0).unsqueeze(1)
    b = (yf - y0).unsqueeze(1)
    xL = x0.long().clamp(0, W - 1)
    xR = (x0.long() + 1).clamp(0, W - 1)
    yT = y0.long().clamp(0, H - 1)
    yB = (y0.long() + 1).clamp(0, H - 1)
    out = (1 - a) * (1 - b) * gather(yT, xL)
    out = out + a * (1 - b) * gather(yT, xR)
    out = out + (1 - a) * b * gather(yB, xL)
    out = out + a * b * gather(yB, xR)
    return out


def resample2d_loops(in1, flow):
    """Scalar-loop numpy statement of ``resample2d`` (bilinear) for tiny inputs."""
    in1 = np.asarray(in1, dtype=np.float64)
    flow = np.asarray(flow, dtype=np.float64)
    B, C, H, W = in1.shape
    _, _, Hf, Wf = flow.shape
    out = np.zeros((B, C, Hf, Wf))
    clampi = lambda v, hi: max(min(v, hi), 0)
    for b in range(B):
        for y in range(Hf):
            for x in range(Wf):
                xf = x + flow[b, 0, y, x]
                yf = y + flow[b, 1, y, x]
                fx, fy = math.floor(xf), math.floor(yf)
                al, be = xf - fx, yf - fy
                xL, xR = clampi(int(fx), W - 1), clampi(int(fx) + 1, W - 1)
                yT, yB = clampi(int(fy), H - 1), clampi(int(fy) + 1, H - 1)
                for c in range(C):
                    im = in1[b, c]
                    out[b, c, y, x] = ((1 - al) * (1 - be) * im[yT, xL] + al * (1 - be) * im[yT, xR]
                                       + (1 - al) * be * im[yB, xL] + al * be * im[yB, xR])
    return out


def channelnorm(x):
    """sqrt(sum_c x^2) -> one channel.  SURVEY.md 8a row a7 (calls flownet2/models.py:141-180)."""
    return torch.sqrt((x * x).sum(dim=1, keepdim=True))


def channelnorm_backward(x, out, grad_out):
    """Upstream backward: gO * x / (out + 1e-8)   [UPSTREAM-RECALLED, SURVEY.md 2.1]."""
    return grad_out * x / (out + 1e-8)


# ---------------------------------------------------------------------------------------------------------------------
# nn.Module forms with the flownet2 constructor signatures, so the tests can put the oracle behind the reference's own
# model classes (tests/test_gpu_reference_models.py) exactly where the CUDA layers go.
class _ChannelNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        out = channelnorm(x)
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, out = ctx.saved_tensors
        return channelnorm_backward(x, out, grad_out)


class Correlation(torch.nn.Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.geom = (pad_size, kernel_size, max_displacement, stride1, stride2)

    def forward(self, input1, input2):
        return correlation(input1, input2, *self.geom)


class Resample2d(torch.nn.Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        assert kernel_size == 1
        self.bilinear = bilinear

    def forward(self, input1, input2):
        return resample2d(input1, input2, self.bilinear)


class ChannelNorm(torch.nn.Module):
    def __init__(self, norm_deg=2):
        super().__init__()

    def forward(self, input1):
        return _ChannelNormFn.apply(input1)


def Correlation_like(layer):
    """The oracle Correlation with the geometry of any object carrying flownet2's attribute names."""
    return Correlation(layer.pad_size, layer.kernel_size, layer.max_displacement, layer.stride1, layer.stride2)


MODULES = {"Correlation": Correlation, "Resample2d": Resample2d, "ChannelNorm": ChannelNorm}

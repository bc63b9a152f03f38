"""CPU oracle for the decoder's flow heads.  TEST INFRASTRUCTURE ONLY.

The reference builds these layers from ``torch.nn`` itself (``PWC/models/PWCNet.py:30-34``)::

    def predict_flow(in_planes): return nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)
    def deconv(in_planes, out_planes, kernel_size=4, stride=2, padding=1):
        return nn.ConvTranspose2d(in_planes, out_planes, kernel_size, stride, padding, bias=True)

so the reference's arithmetic IS PyTorch's convolution: ``predict_flow`` / ``deconv`` below call the same functional
ops in whatever dtype they are given (float64 in the tests) -- PINNED by construction.  ``*_loops`` restate the two
operators as explicit shifted sums, independent of ``torch.nn.functional.conv*``; ``tests/test_oracle.py`` checks the
two statements against each other, which is what fixes the index conventions (cross-correlation, not convolution;
weight layouts ``[out, in, kh, kw]`` / ``[in, out, kh, kw]``; output pixel ``2y - 1 + ky`` of the transposed operator)
that ``csrc/flowhead.cu`` implements.
"""
import torch
import torch.nn.functional as F


def predict_flow(x, weight, bias=None):
    """``nn.Conv2d(C, 2, 3, 1, 1)`` of ``PWCNet.py:30-31`` (``predict_flow6..2`` ``:81-121``, ``dc_conv7`` ``:134``)."""
    return F.conv2d(x, weight, bias, stride=1, padding=1)


def deconv(x, weight, bias=None):
    """``nn.ConvTranspose2d(C, 2, 4, 2, 1)`` of ``PWCNet.py:33-34`` (``deconv6..1``, ``upfeat6..3`` ``:82-124``)."""
    return F.conv_transpose2d(x, weight, bias, stride=2, padding=1)


def predict_flow_loops(x, weight, bias=None):
    """out[b,o,y,x] = bias[o] + sum_c sum_ky sum_kx weight[o,c,ky,kx] * x[b,c,y+ky-1,x+kx-1], zero outside the map."""
    B, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    out = x.new_zeros(B, weight.shape[0], H, W)
    for ky in range(3):
        for kx in range(3):
            out += torch.einsum("oc,bchw->bohw", weight[:, :, ky, kx], xp[:, :, ky:ky + H, kx:kx + W])
    return out if bias is None else out + bias.view(1, -1, 1, 1)


def deconv_loops(x, weight, bias=None):
    """out[b,o,2y-1+ky,2x-1+kx] += x[b,c,y,x] * weight[c,o,ky,kx] for ky, kx in 0..3; out is [B, O, 2H, 2W]."""
    B, C, H, W = x.shape
    big = x.new_zeros(B, weight.shape[1], 2 * H + 2, 2 * W + 2)   # output rows -1 .. 2H, columns -1 .. 2W
    for ky in range(4):
        for kx in range(4):
            big[:, :, ky:ky + 2 * H:2, kx:kx + 2 * W:2] += torch.einsum("co,bchw->bohw", weight[:, :, ky, kx], x)
    out = big[:, :, 1:2 * H + 1, 1:2 * W + 1]
    return out if bias is None else out + bias.view(1, -1, 1, 1)

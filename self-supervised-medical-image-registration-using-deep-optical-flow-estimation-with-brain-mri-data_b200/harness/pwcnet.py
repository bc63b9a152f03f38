"""Host model for the hot path: the reference's 1-channel PWC-DC-Net variant (``PWC/models/PWCNet.py:38-279``).

This is measurement harness, not product: stock cuDNN convolutions laid out as the reference does
(feature pyramid with the extra ``conv*aa`` layers ``:53-70``, DenseNet decoders ``:78-122``, dilated
context network ``:128-134``, 7 flow outputs ``:277``) around the ofr_b200 hot-path layers
(``Correlation(4,1,4,1,1)`` ``:69`` and the fused ``pwc_warp`` replacing ``:143-179``).  It exists
because ``/root/reference`` is not present on the GPU box, so ``bench.py`` cannot import the
reference's class there.  Parameter names and shapes equal the reference's, so a ``state_dict`` moves
between the two unchanged (checked in ``tests/test_harness_cpu.py``: 9 374 184 parameters).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..layers import Correlation, pwc_warp
from ..flowhead import accelerate_flow_heads

# (level, feature channels) from fine to coarse: pyramid levels 1..6
_PYRAMID = [(1, 16), (2, 32), (3, 64), (4, 96), (5, 128), (6, 196)]
_DENSE = [128, 128, 96, 64, 32]               # decoder DenseNet widths
_FLOW_SCALE = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}  # 20 / 2**level  (PWCNet.py:214,228,242,258)


def _conv(cin, cout, stride=1, dilation=1):
    pad = dilation
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, pad, dilation, bias=True), nn.LeakyReLU(0.1))


class PWCNetHost(nn.Module):
    def __init__(self, md=4, corr=None, warp=None, fast_heads=None):
        super().__init__()
        self.md = md
        cin = 1
        for lvl, ch in _PYRAMID:
            # level 6 registers its stride-2 conv as "conv6aa" and the next as "conv6a" (PWCNet.py:68-70)
            first, second = ("aa", "a") if lvl == 6 else ("a", "aa")
            setattr(self, f"conv{lvl}{first}", _conv(cin, ch, stride=2))
            setattr(self, f"conv{lvl}{second}", _conv(ch, ch))
            setattr(self, f"conv{lvl}b", _conv(ch, ch))
            cin = ch
        # product path: LeakyReLU fused into the correlation store, flow pre-scale fused into the warp;
        # injected layers (the CPU oracle in tests / bench reference arm) are followed by the separate ops, as in
        # the reference
        self._fused = corr is None and warp is None
        self.corr = corr if corr is not None else Correlation(pad_size=md, kernel_size=1, max_displacement=md,
                                                              stride1=1, stride2=1, corr_multiply=1, leaky_relu=0.1)
        self._warp = warp if warp is not None else pwc_warp
        self.leakyRELU = nn.LeakyReLU(0.1)
        nd = (2 * md + 1) ** 2
        feat = dict(_PYRAMID)
        for lvl in (6, 5, 4, 3, 2):
            od = nd if lvl == 6 else nd + feat[lvl] + 4
            width = od
            for i, ch in enumerate(_DENSE):
                setattr(self, f"conv{lvl}_{i}", _conv(width, ch))
                width += ch
            setattr(self, f"predict_flow{lvl}", nn.Conv2d(width, 2, 3, 1, 1))
            if lvl > 2:
                setattr(self, f"deconv{lvl}", nn.ConvTranspose2d(2, 2, 4, 2, 1))
                setattr(self, f"upfeat{lvl}", nn.ConvTranspose2d(width, 2, 4, 2, 1))
        self.deconv2 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.deconv1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.deconv0 = nn.ConvTranspose2d(2, 2, 4, 4, 0)  # present but unused in the reference (PWCNet.py:126,274)
        ctx_in = nd + feat[2] + 4 + sum(_DENSE)
        for i, (cout, dil) in enumerate([(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)], start=1):
            setattr(self, f"dc_conv{i}", _conv(ctx_in, cout, dilation=dil))
            ctx_in = cout
        self.dc_conv7 = nn.Conv2d(32, 2, 3, 1, 1)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_in")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        # product path: the two-output-channel layers (predict_flow*, dc_conv7, deconv*, upfeat*) on the flow-head kernels
        # (SURVEY 8f-4); same parameters and state dict.  Off with injected layers (the CPU oracle has no CUDA tensors).
        if self._fused if fast_heads is None else fast_heads:
            accelerate_flow_heads(self)

    def _features(self, im):
        feats, x = {}, im
        for lvl, _ in _PYRAMID:
            first, second = ("aa", "a") if lvl == 6 else ("a", "aa")
            x = getattr(self, f"conv{lvl}{first}")(x)
            x = getattr(self, f"conv{lvl}{second}")(x)
            x = getattr(self, f"conv{lvl}b")(x)
            feats[lvl] = x
        return feats

    def _decode(self, lvl, x):
        for i in range(len(_DENSE)):
            x = torch.cat((getattr(self, f"conv{lvl}_{i}")(x), x), 1)
        return x, getattr(self, f"predict_flow{lvl}")(x)

    def forward(self, x):
        f1 = self._features(x[:, :1])
        f2 = self._features(x[:, 1:])
        flows = {}
        up_flow = up_feat = None
        dense = None
        for lvl in (6, 5, 4, 3, 2):
            if lvl == 6:
                vol = self.corr(f1[6], f2[6])
                inp = vol if self._fused else self.leakyRELU(vol)
            else:
                if self._fused:
                    vol = self.corr(f1[lvl], self._warp(f2[lvl], up_flow, _FLOW_SCALE[lvl]))
                else:
                    vol = self.leakyRELU(self.corr(f1[lvl], self._warp(f2[lvl], up_flow * _FLOW_SCALE[lvl])))
                inp = torch.cat((vol, f1[lvl], up_flow, up_feat), 1)
            dense, flows[lvl] = self._decode(lvl, inp)
            if lvl > 2:
                up_flow = getattr(self, f"deconv{lvl}")(flows[lvl])
                up_feat = getattr(self, f"upfeat{lvl}")(dense)
        ctx = dense
        for i in range(1, 7):
            ctx = getattr(self, f"dc_conv{i}")(ctx)
        flows[2] = flows[2] + self.dc_conv7(ctx)
        flow1 = self.deconv2(flows[2])
        flow0 = self.deconv1(flow1)
        return flow0, flow1, flows[2], flows[3], flows[4], flows[5], flows[6]

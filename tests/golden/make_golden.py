#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the dev container only).

Sources of truth, none of them modified:
  * /root/reference/loss.py            -> loss.npz   (photometric / ncc / smoothness / OFEloss)
  * /root/reference/models.py          -> warps.npz  (opticalFlowReg.stn)
  * /root/reference/PWC/models/PWCNet.py -> warps.npz (PWCDCNet.warp)
  * TVM's numpy ``correlation_nchw_python`` (ships in this image's site-packages) -> corr.npz
  * torch ``F.grid_sample(padding_mode='border', align_corners=True)``           -> resample.npz
  * /root/reference/PWC/models/PWCNet.py ``predict_flow`` / ``deconv`` helpers    -> flowhead.npz (--only-flowhead)

The reference modules need a few absent third-party imports stubbed (monai,
matplotlib, skimage, torchmetrics, RAFT, the custom CUDA layers); the stubs are
empty modules, the reference functions that run are untouched.  Two module
globals are redirected for a GPU-less host: ``loss.device`` (cuda -> cpu) and,
while ``PWCDCNet.warp`` runs, ``Tensor.cuda`` (identity).

    python tests/golden/make_golden.py      # rewrites the .npz files next to this script
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Anything:  # attribute sink for "from x import y"
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, k):
            return _Anything()

    for name in ["monai", "monai.data", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "skimage",
                 "skimage.measure", "torchmetrics", "torchmetrics.functional",
                 "torchmetrics.functional.clustering", "torchmetrics.functional.regression",
                 "RAFT", "RAFT.core", "RAFT.core.raft", "cv2"]:
        m = _stub(name)
        def _ga(k, _A=_Anything):
            if k.startswith("__"):
                raise AttributeError(k)
            return _A()
        m.__getattr__ = _ga  # type: ignore
        m.__path__ = []  # behave as a package
    sys.modules["monai.data"].MetaTensor = torch.Tensor
    for name in ["correlation_package", "correlation_package.correlation",
                 "networks", "networks.resample2d_package", "networks.resample2d_package.resample2d",
                 "networks.channelnorm_package", "networks.channelnorm_package.channelnorm",
                 "flownet2.networks.correlation_package", "flownet2.networks.correlation_package.correlation",
                 "flownet2.networks.resample2d_package", "flownet2.networks.resample2d_package.resample2d",
                 "flownet2.networks.channelnorm_package", "flownet2.networks.channelnorm_package.channelnorm"]:
        m = _stub(name)
        m.__path__ = []
        m.Correlation = m.Resample2d = m.ChannelNorm = torch.nn.Identity


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "PWC", "models"))
    _install_stubs()
    g = torch.Generator().manual_seed(6)  # the reference's training seed (train.py:122)

    # ------------------------------------------------------------------ loss.py
    ref_loss = importlib.import_module("loss")
    ref_loss.device = torch.device("cpu")
    B = 3
    fixed = torch.rand(B, 1, 32, 32, generator=g)
    scales = [(32, 32), (16, 16), (8, 8), (4, 4)]
    flows = [torch.randn(B, 2, h, w, generator=g) * 1.5 for h, w in scales]
    warped = [torch.rand(B, 1, h, w, generator=g) for h, w in scales]
    out = {"fixed": fixed.numpy()}
    for i, (fl, wp) in enumerate(zip(flows, warped)):
        out[f"flow{i}"] = fl.numpy()
        out[f"warped{i}"] = wp.numpy()
        out[f"photo{i}"] = ref_loss.photometric_loss(fixed, wp).numpy()
        out[f"ncc{i}"] = ref_loss.correlation_loss(fixed, wp).numpy()
        out[f"smooth{i}"] = ref_loss.smoothness_loss(fl).numpy()
    p, c, s, tot = ref_loss.OFEloss(flows, warped, fixed)
    out["ofe"] = np.array([p.item(), c.item(), s.item(), tot.item()], dtype=np.float64)
    p, c, s, tot = ref_loss.OFEloss(flows[:2], warped[:2], fixed, lamb_da=0.25, gamma=10.0, zeta=3.0)
    out["ofe_alt"] = np.array([p.item(), c.item(), s.item(), tot.item()], dtype=np.float64)
    # degenerate NCC branch (loss.py:59-60): exactly-representable constant image -> corr = 1 -> loss 0
    const = torch.full((B, 1, 16, 16), 0.5)
    out["ncc_const"] = ref_loss.correlation_loss(fixed, const).numpy()
    out["smooth_zero"] = ref_loss.smoothness_loss(torch.zeros(B, 2, 8, 8)).numpy()
    # gradients of the total w.r.t. warped / flow (autograd through the reference code)
    fl_g = [f.clone().requires_grad_(True) for f in flows]
    wp_g = [w.clone().requires_grad_(True) for w in warped]
    tot = ref_loss.OFEloss(fl_g, wp_g, fixed)[3]
    tot.backward()
    for i in range(len(scales)):
        out[f"gflow{i}"] = fl_g[i].grad.numpy()
        out[f"gwarped{i}"] = wp_g[i].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)

    # ------------------------------------------------------------------ warps
    ref_models = importlib.import_module("models")
    stn = ref_models.opticalFlowReg.stn
    ref_pwc = importlib.import_module("PWCNet")
    out = {}
    frame = torch.rand(2, 1, 24, 24, generator=g)
    for tag, (h, w) in {"a": (24, 24), "b": (12, 12), "c": (6, 10)}.items():
        flow = torch.randn(2, 2, h, w, generator=g) * 2.5
        out[f"stn_flow_{tag}"] = flow.numpy()
        out[f"stn_out_{tag}"] = stn(None, flow, frame).numpy()
    out["stn_frame"] = frame.numpy()
    out["stn_zero"] = stn(None, torch.zeros(1, 2, 8, 8), frame[:1, :, :8, :8].contiguous()).numpy()
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for tag, (C, h, w) in {"a": (5, 16, 16), "b": (3, 9, 13)}.items():
            x = torch.randn(2, C, h, w, generator=g)
            flo = torch.randn(2, 2, h, w, generator=g) * 2.0
            out[f"pwc_x_{tag}"] = x.numpy()
            out[f"pwc_flow_{tag}"] = flo.numpy()
            out[f"pwc_out_{tag}"] = ref_pwc.PWCDCNet.warp(None, x, flo).numpy()
        x = torch.rand(1, 1, 8, 8, generator=g)
        out["pwc_zero_x"] = x.numpy()
        out["pwc_zero_out"] = ref_pwc.PWCDCNet.warp(None, x, torch.zeros(1, 2, 8, 8)).numpy()
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(HERE, "warps.npz"), **out)

    # ------------------------------------------------------------------ correlation (TVM numpy)
    sys.path.insert(0, "/opt/prime-rl/.venv/lib/python3.12/site-packages/tilelang/3rdparty/tvm/python/tvm/topi/testing")
    tvm_corr = importlib.import_module("correlation_nchw_python").correlation_nchw_python
    out = {}
    cases = {  # tag: (B,C,H,W, pad,k,md,s1,s2)
        "pwc": (2, 6, 7, 9, 4, 1, 4, 1, 1),       # PWCNet.py:69
        "fnc": (1, 5, 6, 8, 20, 1, 20, 1, 2),     # FlowNetC.py:31
        "k3": (1, 3, 9, 10, 3, 3, 2, 1, 1),       # general kernel_size (Caffe window convention)
        "s1": (1, 4, 12, 12, 4, 1, 4, 2, 2),      # stride1 > 1 with an exactly divisible extent
    }
    for tag, (B_, C_, H_, W_, pad, k, md, s1, s2) in cases.items():
        a = torch.randn(B_, C_, H_, W_, generator=g).numpy()
        b = torch.randn(B_, C_, H_, W_, generator=g).numpy()
        out[f"{tag}_f1"] = a
        out[f"{tag}_f2"] = b
        out[f"{tag}_params"] = np.array([pad, k, md, s1, s2])
        out[f"{tag}_out"] = tvm_corr(a.astype(np.float64), b.astype(np.float64), k, md, s1, s2, pad, True)
    np.savez_compressed(os.path.join(HERE, "corr.npz"), **out)

    # ------------------------------------------------------------------ resample2d (grid_sample border)
    out = {}
    for tag, (B_, C_, H_, W_) in {"a": (2, 3, 17, 23), "b": (1, 1, 8, 8)}.items():
        img = torch.randn(B_, C_, H_, W_, generator=g, dtype=torch.float64)
        flow = torch.randn(B_, 2, H_, W_, generator=g, dtype=torch.float64) * 6.0
        xs = torch.arange(W_, dtype=torch.float64).view(1, 1, W_) + flow[:, 0]
        ys = torch.arange(H_, dtype=torch.float64).view(1, H_, 1) + flow[:, 1]
        grid = torch.stack((2 * xs / (W_ - 1) - 1, 2 * ys / (H_ - 1) - 1), dim=-1)
        res = F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)
        out[f"{tag}_in1"] = img.numpy().astype(np.float32)
        out[f"{tag}_flow"] = flow.numpy().astype(np.float32)
        # recompute from the float32-rounded inputs so fixture in/out are consistent
        img32 = torch.from_numpy(out[f"{tag}_in1"]).double()
        fl32 = torch.from_numpy(out[f"{tag}_flow"]).double()
        xs = torch.arange(W_, dtype=torch.float64).view(1, 1, W_) + fl32[:, 0]
        ys = torch.arange(H_, dtype=torch.float64).view(1, H_, 1) + fl32[:, 1]
        grid = torch.stack((2 * xs / (W_ - 1) - 1, 2 * ys / (H_ - 1) - 1), dim=-1)
        out[f"{tag}_out"] = F.grid_sample(img32, grid, mode="bilinear", padding_mode="border",
                                          align_corners=True).numpy()
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
    make_regloss(ref_models, ref_loss)
    make_ncc_const(ref_loss)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


def make_regloss(ref_models, ref_loss):
    """regloss.npz: the reference's own chain ``warped = [stn(flow, moving) for flow in flows]`` (models.py:277) ->
    ``OFEloss(flows, warped, fixed)`` (train.py:56), values and d total / d flow by autograd through the reference code.
    Own generator, so the other fixtures do not depend on it."""
    g = torch.Generator().manual_seed(66)
    stn = ref_models.opticalFlowReg.stn
    B = 2
    fixed = torch.rand(B, 1, 32, 32, generator=g)
    moving = torch.rand(B, 1, 32, 32, generator=g)
    scales = [(32, 32), (16, 16), (8, 8), (4, 4)]
    flows = [(torch.randn(B, 2, h, w, generator=g) * 1.5).requires_grad_(True) for h, w in scales]
    warped = [stn(None, fl, moving) for fl in flows]
    p, c, s, tot = ref_loss.OFEloss(flows, warped, fixed)
    tot.backward()
    out = {"fixed": fixed.numpy(), "moving": moving.numpy(),
           "ofe": np.array([p.item(), c.item(), s.item(), tot.item()], dtype=np.float64)}
    for i, (fl, wp) in enumerate(zip(flows, warped)):
        out[f"flow{i}"] = fl.detach().numpy()
        out[f"warped{i}"] = wp.detach().numpy()
        out[f"gflow{i}"] = fl.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "regloss.npz"), **out)


def make_ncc_const(ref_loss):
    """ncc_const.npz: loss.py:52-64 on constant warped images.  The reference short-circuits to corr = 1 only when the
    float32-centred tensor is identically zero (constants with an exact float32 mean); other constants are decided by
    the rounding noise of torch.mean.  Both behaviours are pinned here; ofr_b200 returns 0 for every constant image
    (include/ofr_b200.h, "NCC of a constant image")."""
    g = torch.Generator().manual_seed(67)
    fixed = torch.rand(3, 1, 32, 32, generator=g)
    out = {"fixed": fixed.numpy(), "values": np.array([0.0, 0.5, 0.25, 0.7, 0.3], dtype=np.float32)}
    out["ref_loss"] = np.array([ref_loss.correlation_loss(fixed, torch.full((3, 1, 16, 16), float(v))).item()
                                for v in out["values"]], dtype=np.float64)
    out["centred_is_zero"] = np.array([bool(torch.equal(torch.full((3, 1, 16, 16), float(v)) -
                                                        torch.mean(torch.full((3, 1, 16, 16), float(v))),
                                                        torch.zeros(3, 1, 16, 16))) for v in out["values"]])
    np.savez_compressed(os.path.join(HERE, "ncc_const.npz"), **out)
    return out


def make_flowhead(ref_pwc):
    """flowhead.npz: the reference's own helpers ``predict_flow(in_planes)`` and ``deconv(in_planes, 2)``
    (PWC/models/PWCNet.py:30-34) construct the layers; inputs, parameters, outputs and all gradients of two shapes each,
    float32 on the CPU exactly as the reference module computes them."""
    g = torch.Generator().manual_seed(68)
    out = {}
    for tag, (B, C, H, W) in {"a": (2, 5, 6, 8), "b": (1, 37, 9, 12)}.items():
        for kind, layer in (("conv", ref_pwc.predict_flow(C)), ("deconv", ref_pwc.deconv(C, 2))):
            with torch.no_grad():
                layer.weight.copy_(torch.randn(layer.weight.shape, generator=g) / (3.0 * C ** 0.5))
                layer.bias.copy_(torch.randn(2, generator=g))
            x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
            y = layer(x)
            go = torch.randn(y.shape, generator=g)
            gx, gw, gb = torch.autograd.grad(y, (x, layer.weight, layer.bias), go)
            k = f"{kind}_{tag}_"
            out.update({k + "x": x.detach().numpy(), k + "w": layer.weight.detach().numpy(), k + "b": layer.bias.detach().numpy(),
                        k + "y": y.detach().numpy(), k + "go": go.numpy(), k + "gx": gx.numpy(), k + "gw": gw.numpy(),
                        k + "gb": gb.numpy()})
    np.savez_compressed(os.path.join(HERE, "flowhead.npz"), **out)
    return out


if __name__ == "__main__":
    if "--only-flowhead" in sys.argv:
        sys.path.insert(0, REF)
        sys.path.insert(0, os.path.join(REF, "PWC", "models"))
        _install_stubs()
        o = make_flowhead(importlib.import_module("PWCNet"))
        print("wrote flowhead.npz", sorted(o)[:4], "...")
    elif "--only-ncc-const" in sys.argv:
        sys.path.insert(0, REF)
        _install_stubs()
        _loss = importlib.import_module("loss")
        _loss.device = torch.device("cpu")
        o = make_ncc_const(_loss)
        print("wrote ncc_const.npz", o["values"], o["ref_loss"], o["centred_is_zero"])
    elif "--only-regloss" in sys.argv:  # add the newest fixture without rewriting the others
        sys.path.insert(0, REF)
        sys.path.insert(0, os.path.join(REF, "PWC", "models"))
        _install_stubs()
        _loss = importlib.import_module("loss")
        _loss.device = torch.device("cpu")
        make_regloss(importlib.import_module("models"), _loss)
        print("wrote regloss.npz")
    else:
        main()

"""GPU: the two-output-channel layers of the PWC-Net decoder -- ``predict_flow`` = nn.Conv2d(C, 2, 3, 1, 1)
and ``deconv`` / ``upfeat`` = nn.ConvTranspose2d(C, 2, 4, 2, 1) (PWC/models/PWCNet.py:30-34) -- against the oracle
(``oracle/flowhead_ref.py``: the reference's implementation of these layers IS ``torch.nn``, run here in float64 on the
CPU), values and all gradients."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [(2, 5, 6, 8), (1, 37, 9, 12), (3, 2, 16, 16), (2, 161, 8, 8), (1, 565, 32, 32), (2, 3, 1, 4), (1, 330, 5, 20)]


@pytest.fixture(scope="module")
def ofr():
    import ofr_b200
    assert torch.cuda.is_available()
    ofr_b200._lib.lib()
    return ofr_b200


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _check(name, got, ref, rtol=2e-5):
    ref = ref.to(torch.float64)
    err = (got.detach().cpu().double() - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-6)
    assert err <= rtol * scale, f"{name}: max|err| {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("bias", [True, False])
def test_flow_conv3_matches_torch_float64(ofr, shape, bias):
    B, C, H, W = shape
    m = ofr.FlowConv2d(C, 2, 3, 1, 1, bias=bias)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=_gen(1)) / (3.0 * C ** 0.5))
        if bias:
            m.bias.copy_(torch.randn(2, generator=_gen(2)))
    x = torch.randn(B, C, H, W, generator=_gen(3))
    go = torch.randn(B, 2, H, W, generator=_gen(4))
    xr = x.double().requires_grad_(True)
    wr = m.weight.detach().double().requires_grad_(True)
    br = m.bias.detach().double().requires_grad_(True) if bias else None
    ref = oracle.predict_flow(xr, wr, br)
    grads_ref = torch.autograd.grad(ref, [xr, wr] + ([br] if bias else []), go.double())
    m = m.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    n0 = ofr.launch_count()
    out = m(xd)
    assert ofr.launch_count() > n0, "the layer did not go through the C ABI"
    grads = torch.autograd.grad(out, [xd, m.weight] + ([m.bias] if bias else []), go.to(DEV))
    _check("conv3 out", out, ref)
    for nm, g, r in zip(("gX", "gW", "gB"), grads, grads_ref):
        _check(f"conv3 {nm}", g, r)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("bias", [True, False])
def test_flow_deconv4_matches_torch_float64(ofr, shape, bias):
    B, C, H, W = shape
    m = ofr.FlowUpsample2d(C, 2, 4, 2, 1, bias=bias)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=_gen(5)) / (2.0 * C ** 0.5))
        if bias:
            m.bias.copy_(torch.randn(2, generator=_gen(6)))
    x = torch.randn(B, C, H, W, generator=_gen(7))
    go = torch.randn(B, 2, 2 * H, 2 * W, generator=_gen(8))
    xr = x.double().requires_grad_(True)
    wr = m.weight.detach().double().requires_grad_(True)
    br = m.bias.detach().double().requires_grad_(True) if bias else None
    ref = oracle.deconv(xr, wr, br)
    grads_ref = torch.autograd.grad(ref, [xr, wr] + ([br] if bias else []), go.double())
    m = m.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    n0 = ofr.launch_count()
    out = m(xd)
    assert ofr.launch_count() > n0, "the layer did not go through the C ABI"
    grads = torch.autograd.grad(out, [xd, m.weight] + ([m.bias] if bias else []), go.to(DEV))
    _check("deconv4 out", out, ref)
    for nm, g, r in zip(("gX", "gW", "gB"), grads, grads_ref):
        _check(f"deconv4 {nm}", g, r)


def test_flow_heads_are_reproducible_and_skip_unneeded_gradients(ofr):
    torch.manual_seed(3)
    m = ofr.FlowConv2d(40, 2, 3, 1, 1).to(DEV)
    x = torch.randn(3, 40, 12, 16, device=DEV)
    go = torch.randn(3, 2, 12, 16, device=DEV)
    outs = []
    for _ in range(2):
        xd = x.clone().requires_grad_(True)
        out = m(xd)
        outs.append((out,) + torch.autograd.grad(out, [xd, m.weight, m.bias], go))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    out = m(x)                                # input without gradient: only the weight gradient is computed
    (gw,) = torch.autograd.grad(out, [m.weight], go)
    assert torch.equal(gw, outs[0][2])


def test_layers_outside_the_kernel_shape_use_the_parent_class(ofr):
    """Width not a multiple of 4, other kernel sizes, three output channels: ``nn.Conv2d`` / ``nn.ConvTranspose2d`` as is."""
    torch.manual_seed(4)
    n0 = ofr.launch_count()
    m = ofr.FlowConv2d(7, 2, 3, 1, 1).to(DEV)
    x = torch.randn(2, 7, 6, 10, device=DEV)                      # W % 4 != 0
    assert torch.allclose(m(x), F.conv2d(x, m.weight, m.bias, 1, 1))
    m3 = ofr.FlowConv2d(7, 3, 3, 1, 1).to(DEV)                   # three output channels
    assert m3(torch.randn(2, 7, 8, 8, device=DEV)).shape == (2, 3, 8, 8)
    up = ofr.FlowUpsample2d(2, 2, 4, 4, 0).to(DEV)               # deconv0 of the reference (stride 4): not this kernel
    assert up(torch.randn(1, 2, 8, 8, device=DEV)).shape == (1, 2, 32, 32)
    assert ofr.launch_count() == n0


def test_accelerate_flow_heads_on_the_host_model(ofr):
    """Re-classing the layers of a constructed model keeps its state dict and its outputs (fp32 cuDNN as the yardstick)."""
    from ofr_b200.harness import PWCNetHost, synthetic_pairs
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(6)
        net = PWCNetHost(md=4, fast_heads=False).to(DEV)
        keys = list(net.state_dict().keys())
        x = synthetic_pairs(2, 128, seed=5).to(DEV)
        with torch.no_grad():
            ref = net(x)
        changed = ofr.accelerate_flow_heads(net)
        assert {"predict_flow6", "predict_flow2", "dc_conv7", "deconv6", "upfeat3", "deconv2", "deconv1"} <= set(changed)
        assert "deconv0" not in changed and list(net.state_dict().keys()) == keys
        n0 = ofr.launch_count()
        with torch.no_grad():
            got = net(x)
        assert ofr.launch_count() - n0 >= 9 + 16       # the hot path's 9 launches + at least one per re-classed layer
        for a, b in zip(got, ref):
            assert (a - b).abs().max().item() <= 2e-4 * max(1.0, float(b.abs().max()))
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_accelerate_flow_heads_on_the_reference_pwcdcnet(ofr):
    """The reference's own ``PWCDCNet`` (PWC/models/PWCNet.py, unmodified, from baseline/_ref) with its flow heads
    re-classed after construction: same state dict, same 7 flows in training mode, gradients reach every parameter."""
    from ofr_b200.harness import reference_env, synthetic_pairs
    if reference_env.reference_dir() is None:
        pytest.skip("baseline/_ref not staged")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with reference_env.ReferenceEnv("ofr") as e:
            torch.manual_seed(6)
            net = e.pwcnet().to(DEV)
            net.train()
            keys = list(net.state_dict().keys())
            x = synthetic_pairs(2, 128, seed=9).to(DEV)
            with torch.no_grad():
                ref = net(x)
            changed = ofr.accelerate_flow_heads(net)
            assert len(changed) == 5 + 1 + 6 + 4 and "deconv0" not in changed   # predict_flow6..2, dc_conv7, deconv6..1, upfeat6..3
            assert list(net.state_dict().keys()) == keys
            n0 = ofr.launch_count()
            got = net(x)
            assert ofr.launch_count() - n0 >= 9 + 16
            assert len(got) == len(ref) == 7
            for a, b in zip(got, ref):
                assert (a - b).abs().max().item() <= 2e-4 * max(1.0, float(b.abs().max()))
            sum(f.square().mean() for f in got).backward()
            missing = [n for n, p_ in net.named_parameters() if p_.grad is None and not n.startswith("deconv0")]
            assert not missing, missing
            for n, p_ in net.named_parameters():
                if p_.grad is not None:
                    assert bool(torch.isfinite(p_.grad).all()), n
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("kind,tag", [("conv", "a"), ("conv", "b"), ("deconv", "a"), ("deconv", "b")])
def test_flow_heads_vs_reference_golden(ofr, golden, kind, tag):
    """The kernels against what the reference's own ``predict_flow`` / ``deconv`` layers computed (flowhead.npz)."""
    g = golden("flowhead")
    k = f"{kind}_{tag}_"
    x, w, b, go = (torch.from_numpy(g[k + n]).to(DEV) for n in ("x", "w", "b", "go"))
    C = x.shape[1]
    m = (ofr.FlowConv2d(C, 2, 3, 1, 1) if kind == "conv" else ofr.FlowUpsample2d(C, 2, 4, 2, 1)).to(DEV)
    with torch.no_grad():
        m.weight.copy_(w)
        m.bias.copy_(b)
    xd = x.clone().requires_grad_(True)
    n0 = ofr.launch_count()
    y = m(xd)
    assert ofr.launch_count() > n0
    gx, gw, gb = torch.autograd.grad(y, (xd, m.weight, m.bias), go)
    for name, got, ref in (("y", y, g[k + "y"]), ("gx", gx, g[k + "gx"]), ("gw", gw, g[k + "gw"]), ("gb", gb, g[k + "gb"])):
        _check(f"{kind}_{tag} {name}", got, torch.from_numpy(ref), rtol=2e-5)

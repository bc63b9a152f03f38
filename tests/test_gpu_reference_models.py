"""GPU, in situ: the reference's OWN model / loss files, unmodified, running on the ofr_b200 layers.

The files are imported from ``baseline/_ref/`` (staged byte for byte from /root/reference by
``tools/stage_reference.py``; see ``harness/reference_env.py``).  The model tests build the reference class with the same
weights on the CUDA layers at the import paths the reference expects (``install_shims()``), on the oracle's layers in
float32 and float64 (plain PyTorch, same device) and on (oracle forward values, CUDA backward); they replay every
hot-path call in isolation on the tensors it saw inside the network, compare flows against the noise floor the two
oracle networks define, and weight gradients on the oracle's own forward trajectory.  TF32 is off and cuDNN deterministic, so what differs between the networks is only the hot path.

Call sites exercised: ``PWC/models/PWCNet.py:69,200-259`` (Correlation(4,1,4,1,1) at five pyramid levels),
``flownet2/networks/FlowNetC.py:31,88`` (Correlation(20,1,20,1,2), C=256), ``flownet2/models.py:40,52-88,121-191``
(4x Resample2d, 6x ChannelNorm), ``loss.py:66-84`` (OFEloss) and ``models.py:256-268`` (stn).
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def env():
    import ofr_b200
    from ofr_b200.harness import reference_env
    assert torch.cuda.is_available()
    ofr_b200._lib.lib()
    if reference_env.reference_dir() is None:
        pytest.skip("baseline/_ref not staged (python tools/stage_reference.py)")
    return reference_env


@pytest.fixture()
def fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _pairs(B, S, seed=6):
    from ofr_b200.harness import synthetic_pairs
    return synthetic_pairs(B, S, seed=seed).to(DEV)


class _AsFloat64(torch.nn.Module):
    """An oracle layer evaluated in float64 and rounded back: differs from the float32 oracle by last-bit rounding only.
    The distance between the two oracle networks is the noise floor of an in-situ comparison -- these networks amplify
    last-bit differences of the cost volume (bilinear-warp cell crossings, LeakyReLU kinks, the 0.9999 mask threshold)."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, *xs):
        return self.inner(*[x.double() for x in xs]).float()


def _o64(name):
    cls = oracle.MODULES[name]
    return lambda *a, **k: _AsFloat64(cls(*a, **k))


ORACLE64 = {"Correlation": _o64("Correlation"), "Resample2d": _o64("Resample2d"), "ChannelNorm": _o64("ChannelNorm")}


class _OracleValuesCudaBackward(torch.nn.Module):
    """Forward VALUES of the float32 oracle, backward through the CUDA layer: ``o + (ref - o)`` is exactly ``ref`` in
    float32 (Sterbenz), and autograd only sees the CUDA call.  A network built on these layers follows the oracle
    network's forward trajectory bit for bit, so its weight gradients differ from the oracle network's by the rounding
    of the CUDA backward kernels alone -- no last-bit forward difference for the network to amplify."""

    def __init__(self, cuda_layer, oracle_layer):
        super().__init__()
        self.cuda_layer, self.oracle_layer = cuda_layer, oracle_layer

    def forward(self, *xs):
        o = self.cuda_layer(*xs)
        with torch.no_grad():
            d = self.oracle_layer(*xs) - o
        return o + d


def _hybrid(name):
    import ofr_b200
    return lambda *a, **k: _OracleValuesCudaBackward(getattr(ofr_b200, name)(*a, **k), oracle.MODULES[name](*a, **k))


def _build_nets(env, ctor):
    """The reference class four times with the same weights: on the CUDA layers, on the float32 oracle, on the float64
    oracle, and on (oracle forward values, CUDA backward)."""
    import ofr_b200
    with env.ReferenceEnv("ofr") as e:
        torch.manual_seed(6)
        net = getattr(e, ctor)().to(DEV)
        kinds = {type(m) for m in net.modules()}
        assert ofr_b200.Correlation in kinds, "the reference class did not pick up ofr_b200.Correlation"
    nets = [net]
    hybrid = {n: _hybrid(n) for n in ("Correlation", "Resample2d", "ChannelNorm")}
    for layers in (oracle.MODULES, ORACLE64, hybrid):
        with env.ReferenceEnv(layers) as e:
            other = getattr(e, ctor)().to(DEV)
        if layers is not hybrid:
            assert ofr_b200.Correlation not in {type(m) for m in other.modules()}
        other.load_state_dict(net.state_dict())
        nets.append(other)
    return nets, kinds


def _capture_calls(net, kinds):
    """Forward hooks on every hot-path layer of ``net``: records (module, inputs, output) per call and keeps the
    output's gradient, so each call can be replayed in isolation on exactly the tensors it saw inside the network."""
    calls, handles = [], []

    def hook(mod, inp, out):
        calls.append((mod, [t.detach().clone() for t in inp], out))
        if out.requires_grad:
            out.retain_grad()

    for m in net.modules():
        if type(m) in kinds:
            handles.append(m.register_forward_hook(hook))
    return calls, handles


def _check_calls_in_isolation(calls, ofr_b200):
    """Every captured call again, alone: the CUDA layer against the oracle on the SAME inputs and upstream gradient
    (float64 oracle for Correlation / ChannelNorm; Resample2d in the layer's own float32 position arithmetic).
    North-star tolerances: 1e-5 max-abs forward, 1e-4 of the gradient's magnitude backward."""
    seen = set()
    for mod, inps, out in calls:
        go = out.grad if out.grad is not None else torch.ones_like(out)
        name = type(mod).__name__
        seen.add(name)
        xs = [t.clone().requires_grad_(True) for t in inps]
        got = mod(*xs)
        gg = torch.autograd.grad(got, xs, go, allow_unused=True)
        if name == "Resample2d":
            rs = [t.cpu().requires_grad_(True) for t in inps]
            ref = oracle.MODULES[name]()(*rs)
            rg = torch.autograd.grad(ref, rs, go.cpu(), allow_unused=True)
        else:
            rs = [t.double().requires_grad_(True) for t in inps]
            ref = (oracle.Correlation_like(mod) if name == "Correlation" else oracle.MODULES[name]())(*rs)
            rg = torch.autograd.grad(ref, rs, go.double(), allow_unused=True)
        scale = max(1.0, float(ref.abs().max()))
        err = float((got.detach().double().cpu() - ref.detach().double().cpu()).abs().max())
        assert err <= 1e-5 * scale, (name, tuple(inps[0].shape), "fwd", err)
        for i, (a, b) in enumerate(zip(gg, rg)):
            if b is None:
                continue
            assert _rel(a.cpu(), b.cpu()) <= 1e-4, (name, tuple(inps[0].shape), f"grad[{i}]", _rel(a.cpu(), b.cpu()))
    return seen


def _run_all(nets, kinds, x, cot_seed):
    """forward + backward of the networks with one shared cotangent; returns flows, grads and the captured
    hot-path calls of the CUDA-layer network."""
    g = torch.Generator(device=DEV).manual_seed(cot_seed)
    results, cot, calls = [], None, None
    for i, net in enumerate(nets):
        if i == 0:
            calls, handles = _capture_calls(net, kinds)
        out = net(x)
        flows = list(dict.fromkeys(out))  # FlowNet2 returns the same tensor twice
        if i == 0:
            for h in handles:
                h.remove()
            cot = [torch.randn(f.shape, device=DEV, generator=g) for f in flows]
        sum((f * c).sum() for f, c in zip(flows, cot)).backward()
        results.append(([f.detach() for f in flows], {n: p.grad for n, p in net.named_parameters() if p.grad is not None}))
    return results, calls


def _l2_rel(ga, gb):
    num = sum(float((ga[n].double() - gb[n].double()).pow(2).sum()) for n in gb)
    den = sum(float(gb[n].double().pow(2).sum()) for n in gb)
    return (num / den) ** 0.5


def _assert_in_situ(results, what):
    """Forward: flows of the CUDA network within 3x the distance between the float32- and float64-oracle networks (the
    noise floor of the comparison).  Backward: the (oracle forward values, CUDA backward) network against the float32
    oracle network -- same forward trajectory, so the weight gradients must agree to the north-star gradient tolerance.
    The weight gradients of the pure CUDA network are printed with the floor: the networks amplify last-bit FORWARD
    differences of the cost volume by ~1e4 (measured: the floor itself moves between 2e-4 and 4e-3 with the input seed
    and the kind of last-bit perturbation), so that number says nothing about the backward kernels."""
    (f_c, g_c), (f_32, g_32), (f_64, g_64), (f_h, g_h) = results
    floor_f = max(_rel(a, b) for a, b in zip(f_32, f_64))
    err_f = max(_rel(a, b) for a, b in zip(f_c, f_64))
    assert err_f <= 3 * floor_f + 1e-5, (what, "flows", err_f, floor_f)
    for a, b in zip(f_h, f_32):
        assert torch.equal(a, b), (what, "the hybrid network left the oracle's forward trajectory")
    assert set(g_c) == set(g_h) == set(g_32)
    err_h = _l2_rel(g_h, g_32)
    gmax = max(float(g_32[n].abs().max()) for n in g_32)
    # per-tensor check on the tensors that carry gradient (a convolution bias in front of a BatchNorm has none: noise)
    worst_h = max(_rel(g_h[n], g_32[n]) for n in g_32 if float(g_32[n].abs().max()) > 1e-4 * gmax)
    print(f"{what}: flows err {err_f:.2e} (floor {floor_f:.2e}); weight gradients, CUDA backward on the oracle's trajectory: "
          f"L2 {err_h:.2e}, worst tensor {worst_h:.2e}; pure CUDA network L2 {_l2_rel(g_c, g_64):.2e} "
          f"(floor {_l2_rel(g_32, g_64):.2e})")
    assert err_h <= 1e-4 and worst_h <= 1e-3, (what, "weight gradients", err_h, worst_h)


@pytest.fixture()
def deterministic_cudnn():
    old = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    yield
    torch.backends.cudnn.deterministic = old


def test_reference_pwcnet_unmodified_on_cuda_layers(env, fp32_convs, deterministic_cudnn):
    """``PWC/models/PWCNet.py`` PWCDCNet(md=4), B=2, 256x256, forward + backward (the reference's own ``warp`` method,
    grid_sample, runs on all sides).  (i) each of the five Correlation calls replayed in isolation on the tensors it
    saw inside the network, against the float64 oracle, at the north-star tolerances; (ii) the 7 flows of the network
    on the CUDA layers within 3x the noise floor of this comparison (distance between the same network on the float32
    and on the float64 oracle); (iii) ALL weight gradients with the CUDA backward kernels on the oracle's forward
    trajectory against the oracle network (``_assert_in_situ``)."""
    import ofr_b200
    nets, kinds = _build_nets(env, "pwcnet")
    assert sum(p.numel() for p in nets[0].parameters()) == 9374184
    x = _pairs(2, 256)
    n0 = ofr_b200.launch_count()
    results, calls = _run_all(nets, {ofr_b200.Correlation}, x, 7)
    assert ofr_b200.launch_count() - n0 >= 10, "five Correlation forwards + backwards expected (PWCNet.py:200-259)"
    assert [f.shape[-1] for f in results[0][0]] == [256, 128, 64, 32, 16, 8, 4]
    assert [tuple(c[1][0].shape[1:]) for c in calls] == [(196, 4, 4), (128, 8, 8), (96, 16, 16), (64, 32, 32), (32, 64, 64)]
    assert _check_calls_in_isolation(calls, ofr_b200) == {"Correlation"}
    _assert_in_situ(results, "PWCDCNet")


def test_reference_flownet2_unmodified_on_cuda_layers(env, fp32_convs, deterministic_cudnn):
    """``flownet2/models.py`` FlowNet2(batchNorm=True), B=2, 256x256: the FlowNetC correlation (C=256, 32x32 map: the
    tcgen05 kernels), 4 Resample2d and 6 ChannelNorm calls per forward (models.py:139-180), forward and backward; same
    three checks as the PWC-Net test."""
    import ofr_b200
    nets, kinds = _build_nets(env, "flownet2")
    assert ofr_b200.Resample2d in kinds and ofr_b200.ChannelNorm in kinds
    assert sum(p.numel() for p in nets[0].parameters()) == 162493554
    x = _pairs(2, 256)
    n0 = ofr_b200.launch_count()
    results, calls = _run_all(nets, {ofr_b200.Correlation, ofr_b200.Resample2d, ofr_b200.ChannelNorm}, x, 8)
    assert ofr_b200.launch_count() - n0 >= 11
    names = [type(c[0]).__name__ for c in calls]
    assert names.count("Correlation") == 1 and names.count("Resample2d") == 4 and names.count("ChannelNorm") == 6
    assert tuple(calls[0][1][0].shape) == (2, 256, 32, 32)               # FlowNetC.py:88
    assert results[0][0][0].shape == (2, 2, 256, 256)
    assert _check_calls_in_isolation(calls, ofr_b200) == {"Correlation", "Resample2d", "ChannelNorm"}
    _assert_in_situ(results, "FlowNet2")


def test_reference_loss_and_stn_unmodified_vs_fused_kernels(env):
    """``loss.OFEloss`` (loss.py:66-84) and ``opticalFlowReg.stn`` (models.py:256-268) from the reference's files, run
    on the GPU as train.py:50-56 chains them, against ``ofr_b200.registration_loss`` on the same flows."""
    import ofr_b200
    with env.ReferenceEnv("ofr") as e:
        ref_loss = e.loss_module()
        ref_models = e.models_module()
        stn = ref_models.opticalFlowReg.stn
        B, S = 3, 64
        g = torch.Generator(device=DEV).manual_seed(9)
        pair = _pairs(B, S, seed=9)
        fixed, moving = pair[:, :1].contiguous(), pair[:, 1:].contiguous()
        flows = [(torch.randn(B, 2, S >> i, S >> i, device=DEV, generator=g) * 1.5) for i in range(4)]
        fr = [f.clone().requires_grad_(True) for f in flows]
        warped_r = [stn(None, f, moving) for f in fr]
        terms_r = ref_loss.OFEloss(fr, warped_r, fixed)
        grads_r = torch.autograd.grad(terms_r[3], fr)
    fo = [f.clone().requires_grad_(True) for f in flows]
    terms, warped = ofr_b200.registration_loss(fo, moving, fixed)
    grads = torch.autograd.grad(terms[3], fo)
    np.testing.assert_allclose([float(t) for t in terms], [float(t) for t in terms_r], rtol=5e-6)
    for i in range(4):
        assert (warped[i] - warped_r[i].detach()).abs().max().item() <= 1e-5
        h, w = flows[i].shape[2:]
        diff = torch.nn.functional.interpolate(fixed, (h, w), mode="bilinear", align_corners=False) - warped[i]
        keep = (diff.abs() > 1e-3).float()      # the |x|^-0.5 photometric derivative is ill-conditioned near 0
        err = ((grads[i] - grads_r[i]) * keep).abs().max().item()
        assert err <= 1e-6 + 1e-4 * (grads_r[i] * keep).abs().max().item(), (i, err)


def test_reference_training_chain_pwcnet_stn_ofeloss(env, fp32_convs, deterministic_cudnn):
    """train.py:50-57 end to end with the reference's files: PWCDCNet (CUDA Correlation) -> ``opticalFlowReg.stn`` per
    scale -> ``loss.OFEloss`` -> backward, against the product chain (same net -> ``registration_loss``): the four loss
    terms, and d total / d flow of every scale away from the pixels where the |x|^-0.5 photometric derivative is
    ill-conditioned in float32 (on registered pairs those pixels dominate the weight gradients of BOTH chains, which is
    why weight gradients are not compared here)."""
    import ofr_b200
    with env.ReferenceEnv("ofr") as e:
        torch.manual_seed(6)
        net = e.pwcnet().to(DEV)
        ref_loss = e.loss_module()
        stn = e.models_module().opticalFlowReg.stn
        x = _pairs(2, 256, seed=11)
        fixed, moving = x[:, :1].contiguous(), x[:, 1:].contiguous()
        flows = [f.detach().requires_grad_(True) for f in net(x)]
        warped_r = [stn(None, f, moving) for f in flows]
        terms_r = ref_loss.OFEloss(flows, warped_r, fixed)
        grads_r = torch.autograd.grad(terms_r[3], flows)
    flows2 = [f.detach().requires_grad_(True) for f in flows]
    terms, warped = ofr_b200.registration_loss(flows2, moving, fixed)
    grads = torch.autograd.grad(terms[3], flows2)
    np.testing.assert_allclose([float(t) for t in terms], [float(t) for t in terms_r], rtol=1e-5)
    for i in range(len(flows)):
        assert (warped[i] - warped_r[i].detach()).abs().max().item() <= 1e-5
        h, w = flows[i].shape[2:]
        diff = torch.nn.functional.interpolate(fixed, (h, w), mode="bilinear", align_corners=False) - warped[i]
        keep = (diff.abs() > 1e-3).float()
        err = ((grads[i] - grads_r[i]) * keep).abs().max().item()
        assert err <= 1e-6 + 1e-4 * (grads_r[i] * keep).abs().max().item(), (i, err)

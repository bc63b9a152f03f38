"""GPU, in situ: the reference's OWN model / loss files, unmodified, running on the ofr_b200 layers.

The files are imported from ``baseline/_ref/`` (staged byte for byte from /root/reference by
``tools/stage_reference.py``; see ``harness/reference_env.py``).  Each test builds the reference class twice with the
same weights -- once with the CUDA layers at the import paths the reference expects (``install_shims()``), once with
the oracle's layers (plain PyTorch, same device, float32) -- and compares flows and weight gradients.  TF32 is
switched off so the convolutions on both sides are the same float32 arithmetic; what differs is only the hot path.

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


def _build_twice(env, ctor):
    """(reference class on the CUDA layers, the same class on the oracle layers, same weights)."""
    import ofr_b200
    with env.ReferenceEnv("ofr") as e:
        torch.manual_seed(6)
        net = getattr(e, ctor)().to(DEV)
        kinds = {type(m) for m in net.modules()}
        assert ofr_b200.Correlation in kinds, "the reference class did not pick up ofr_b200.Correlation"
    with env.ReferenceEnv(oracle.MODULES) as e:
        net_o = getattr(e, ctor)().to(DEV)
        kinds_o = {type(m) for m in net_o.modules()}
        assert oracle.MODULES["Correlation"] in kinds_o and ofr_b200.Correlation not in kinds_o
    net_o.load_state_dict(net.state_dict())
    return net, net_o, kinds


def test_reference_pwcnet_unmodified_on_cuda_layers(env, fp32_convs):
    """``PWC/models/PWCNet.py`` PWCDCNet(md=4), B=2, 256x256: 7 flows and weight gradients, CUDA layers vs oracle
    layers behind the same class (the reference's own ``warp`` method, grid_sample, runs on both sides)."""
    import ofr_b200
    net, net_o, _ = _build_twice(env, "pwcnet")
    assert sum(p.numel() for p in net.parameters()) == 9374184
    x = _pairs(2, 256)
    n0 = ofr_b200.launch_count()
    flows = net(x)
    assert ofr_b200.launch_count() - n0 >= 5, "five Correlation calls expected (PWCNet.py:200-259)"
    flows_o = net_o(x)
    assert len(flows) == 7 and [f.shape[-1] for f in flows] == [256, 128, 64, 32, 16, 8, 4]
    for i, (a, b) in enumerate(zip(flows, flows_o)):
        assert _rel(a, b) < 1e-4, (i, _rel(a, b))
    g = torch.Generator(device=DEV).manual_seed(7)
    cot = [torch.randn(f.shape, device=DEV, generator=g) for f in flows]
    sum((f * c).sum() for f, c in zip(flows, cot)).backward()
    sum((f * c).sum() for f, c in zip(flows_o, cot)).backward()
    po = dict(net_o.named_parameters())
    worst = 0.0
    for name, p in net.named_parameters():
        if p.grad is None:
            assert po[name].grad is None
            continue
        worst = max(worst, _rel(p.grad, po[name].grad))
    assert worst < 2e-3, worst


def test_reference_flownet2_unmodified_on_cuda_layers(env, fp32_convs):
    """``flownet2/models.py`` FlowNet2(batchNorm=True), B=2, 256x256: the FlowNetC correlation (C=256, 32x32 map:
    the tcgen05 kernels), 4 Resample2d and 6 ChannelNorm calls per forward, forward and backward."""
    import ofr_b200
    net, net_o, kinds = _build_twice(env, "flownet2")
    assert ofr_b200.Resample2d in kinds and ofr_b200.ChannelNorm in kinds
    assert sum(p.numel() for p in net.parameters()) == 162493554
    x = _pairs(2, 256)
    n0 = ofr_b200.launch_count()
    out = net(x)
    launched = ofr_b200.launch_count() - n0
    assert launched >= 11, launched                                      # 1 corr + 4 resample + 6 chnorm
    out_o = net_o(x)
    assert out[0].shape == (2, 2, 256, 256) and out[0] is out[1]         # flownet2/models.py:191
    assert _rel(out[0], out_o[0]) < 1e-3, _rel(out[0], out_o[0])
    g = torch.Generator(device=DEV).manual_seed(8)
    cot = torch.randn(out[0].shape, device=DEV, generator=g)
    (out[0] * cot).sum().backward()
    (out_o[0] * cot).sum().backward()
    po = dict(net_o.named_parameters())
    checked = 0
    for name in ["flownetc.conv1.0.weight", "flownetc.conv3.0.weight", "flownetc.conv_redir.0.weight",
                 "flownets_1.conv1.0.weight", "flownets_2.conv1.0.weight", "flownets_d.conv0.0.weight",
                 "flownetfusion.conv0.0.weight", "flownetfusion.predict_flow0.weight"]:
        p = dict(net.named_parameters())[name]
        r = _rel(p.grad, po[name].grad)
        assert r < 5e-3, (name, r)
        checked += 1
    assert checked == 8


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


def test_reference_training_chain_pwcnet_stn_ofeloss(env, fp32_convs):
    """train.py:50-57 end to end with the reference's files: PWCDCNet (CUDA Correlation) -> stn per scale -> OFEloss
    -> backward, against the product chain (same net -> ``registration_loss``): loss terms and weight gradients."""
    import ofr_b200
    with env.ReferenceEnv("ofr") as e:
        torch.manual_seed(6)
        net = e.pwcnet().to(DEV)
        ref_loss = e.loss_module()
        stn = e.models_module().opticalFlowReg.stn
        x = _pairs(2, 256, seed=11)
        fixed, moving = x[:, :1].contiguous(), x[:, 1:].contiguous()
        flows = net(x)
        terms_r = ref_loss.OFEloss(flows, [stn(None, f, moving) for f in flows], fixed)
        terms_r[3].backward()
        grads_r = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        net.zero_grad(set_to_none=True)
        flows = net(x)
        terms, _ = ofr_b200.registration_loss(list(flows), moving, fixed)
        terms[3].backward()
    np.testing.assert_allclose([float(t) for t in terms], [float(t) for t in terms_r], rtol=1e-5)
    worst = max(_rel(p.grad, grads_r[n]) for n, p in net.named_parameters() if p.grad is not None)
    assert worst < 5e-3, worst

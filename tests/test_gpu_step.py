"""GPU: the pieces around the layers that SURVEY.md section 8f ranks next -- the de-synchronised ``opticalFlowReg``
wrapper (8f-2), the training step captured as one CUDA graph (8f-2) and the on-device pair generator (8f-3)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ofr():
    import ofr_b200
    assert torch.cuda.is_available()
    ofr_b200._lib.lib()
    return ofr_b200


def test_optical_flow_reg_wrapper_vs_reference_forward(ofr):
    """``ofr_b200.OpticalFlowReg`` against the reference's own ``opticalFlowReg.forward`` (models.py:270-289, imported
    unmodified from baseline/_ref; its ``__init__`` is bypassed because it loads a checkpoint from a C:\\ path,
    SURVEY.md appendix A) around the same predictor: flows, the 7 warped images, the rounded / clamped warped labels
    (the reference takes them through numpy on the host, ``:286``) and the warped line grid."""
    from ofr_b200.harness import reference_env, PWCNetHost, synthetic_pairs
    if reference_env.reference_dir() is None:
        pytest.skip("baseline/_ref not staged")
    torch.manual_seed(6)
    net = PWCNetHost(md=4).to(DEV)
    # batch 1: the reference's forward gives grid_sample a batch-1 grid image with the batch's flow (models.py:279-281)
    # and therefore only runs at B = 1; the wrapper itself takes any batch (checked below)
    x = synthetic_pairs(1, 256, seed=21).to(DEV)
    segs = torch.randint(0, 4, (1, 2, 256, 256), generator=torch.Generator().manual_seed(22)).float().to(DEV)
    with reference_env.ReferenceEnv("ofr") as e:
        models = e.models_module()
        ref = models.opticalFlowReg.__new__(models.opticalFlowReg)
        torch.nn.Module.__init__(ref)
        ref.predictor = net
        with torch.no_grad():
            rf, rw, rs, rg = ref(x, segs)
    with torch.no_grad():
        f, w, s, g = ofr.OpticalFlowReg(net)(x, segs)
    assert len(f) == len(rf) == 7 and len(w) == len(rw) == 7
    for a, b in zip(f, rf):   # same predictor, two passes (cuDNN's default algorithms are not bitwise reproducible)
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, float(b.abs().max()))
    for i, (a, b) in enumerate(zip(ofr.stn_warp_multi(list(rf), x[:, 1:2].contiguous()), rw)):
        assert (a - b).abs().max().item() <= 1e-5, i      # the kernels on the reference pass's own flows
    for i, (a, b) in enumerate(zip(w, rw)):
        assert (a - b).abs().max().item() <= 1e-4, i      # the wrapper's own pass (flows differ in the last bits)
    assert s.shape == rs.shape == (1, 1, 256, 256) and g.shape == rg.shape
    assert (g - rg).abs().max().item() <= 1e-4
    # labels: a sample within 1e-5 of a .5 boundary may round the other way; everything else must agree exactly
    pre = oracle.stn_warp(f[0].cpu(), segs[:, 1:2].cpu())
    safe = ((pre - pre.floor() - 0.5).abs() > 1e-4).to(DEV)
    assert torch.equal(s[safe], rs.to(s.dtype)[safe])
    assert float(s.min()) >= 0 and float(s.max()) <= 3 and torch.equal(s, s.round())
    xb = synthetic_pairs(3, 256, seed=23).to(DEV)
    with torch.no_grad():
        fb, wb, sb, gb = ofr.OpticalFlowReg(net)(xb, segs.expand(3, 2, 256, 256).contiguous())
    assert sb.shape == gb.shape == (3, 1, 256, 256) and len(wb) == 7


def test_training_step_as_one_cuda_graph_matches_eager(ofr):
    """Forward + fused loss + backward + Adam captured in ONE CUDA graph (no host sync, no allocation inside the C ABI):
    replaying it from the same initial weights reproduces the eager loss trajectory."""
    from ofr_b200.harness import PWCNetHost, RegistrationStep, synthetic_pairs, make_optimizer, train_step, GraphedStep
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = False

    def build():
        torch.manual_seed(6)
        host = PWCNetHost(md=4).to(DEV)
        host.deconv0.requires_grad_(False)
        step = RegistrationStep(host, ofr.stn_warp, ofr.OFEloss, reg_loss=ofr.registration_loss)
        return step, make_optimizer(host, capturable=True)

    try:
        pair = synthetic_pairs(2, 128, seed=6).to(DEV)
        step, opt = build()
        n0 = ofr.launch_count()
        graphed = GraphedStep(step, opt, pair, warm_iters=3)   # 3 eager warm-up steps, then the capture pass
        assert ofr.launch_count() - n0 >= 4 * 20              # the hot-path launches were recorded into the graph too
        params = [p for p in step.predictor.parameters() if p.requires_grad]
        state = [t for p in params for t in opt.state[p].values() if torch.is_tensor(t)]
        snap = [t.clone() for t in params + state]

        def restore():
            with torch.no_grad():
                for t, v in zip(params + state, snap):
                    t.copy_(v)

        # one step from the SAME weights and optimizer state, eagerly and as a graph replay
        loss_e = float(train_step(step, opt, pair))
        after_e = [p.detach().clone() for p in params]
        restore()
        loss_g = float(graphed.replay())
        after_g = [p.detach().clone() for p in params]
        assert abs(loss_g - loss_e) <= 1e-5 * abs(loss_e), (loss_g, loss_e)
        moved = sum(float((a - s_).abs().sum()) for a, s_ in zip(after_g, snap))
        assert moved > 0                                       # the replay really stepped the optimizer
        # Adam's first steps move every weight by ~lr whatever the gradient's size, so a gradient component whose sign
        # is decided by rounding noise flips its update: compare the bulk, not the worst element
        close_frac = sum(int(((a - b).abs() <= 2e-6).sum()) for a, b in zip(after_g, after_e)) / sum(p.numel() for p in params)
        assert close_frac > 0.98, close_frac
        # replays from the same state agree (cuDNN's default algorithms are not bitwise reproducible)
        restore()
        again = float(graphed.replay())
        assert abs(again - loss_g) <= 1e-6 * abs(loss_g)
    finally:
        torch.backends.cudnn.benchmark = old


def test_elastic_pair_generator_on_device(ofr):
    """pairs.py (replaces Rand2DElasticd + ScaleIntensityd of dataset.py:72-85 on the device): shapes, [0,1] range,
    seed reproducibility, displacement statistics of spacing-16 / magnitude (0,0.5) control points, and the warp
    against the oracle's Resample2d."""
    gen = ofr.ElasticPairGenerator(size=256, device=DEV, seed=6)
    pair = gen(8)
    assert pair.shape == (8, 2, 256, 256) and pair.is_cuda and pair.dtype == torch.float32
    assert float(pair.min()) == 0.0 and float(pair.max()) == 1.0
    assert float((pair[:, 0] - pair[:, 1]).abs().max()) > 1e-3
    again = ofr.ElasticPairGenerator(size=256, device=DEV, seed=6)(8)
    assert torch.equal(pair, again)
    d = ofr.ElasticPairGenerator(size=256, device=DEV, seed=7, magnitude_range=(0.5, 1.0)).displacement(16)
    assert d.shape == (16, 2, 256, 256)
    assert 0.1 < float(d.std()) < 1.5 and float(d.abs().max()) < 8.0
    # the deformation itself: Resample2d kernel vs the oracle on the same field
    g2 = ofr.ElasticPairGenerator(size=64, device=DEV, seed=8, magnitude_range=(1.0, 2.0))
    fixed = torch.rand(2, 1, 64, 64, device=DEV)
    flow = g2.displacement(2)
    g2.gen.manual_seed(8)
    out, labels = g2.deform(fixed, segs=torch.randint(0, 4, (2, 1, 64, 64), device=DEV).float())
    ref = oracle.resample2d(fixed.cpu(), flow.cpu())
    both = torch.cat((fixed.cpu(), ref), 1)
    lo, hi = both.amin(dim=(1, 2, 3), keepdim=True), both.amax(dim=(1, 2, 3), keepdim=True)
    assert ((both - lo) / (hi - lo) - out.cpu()).abs().max().item() <= 1e-5
    assert labels.shape == (2, 2, 64, 64) and torch.equal(labels, labels.round())

"""CPU: pin the oracle against the golden fixtures (reference-generated) and known answers."""
import numpy as np
import pytest
import torch

import oracle


def T(a, dt=torch.float64):
    return torch.from_numpy(np.asarray(a)).to(dt)


# ----------------------------------------------------------------------------- correlation
@pytest.mark.parametrize("tag", ["pwc", "fnc", "k3", "s1"])
def test_correlation_vs_tvm_golden(golden, tag):
    g = golden("corr")
    pad, k, md, s1, s2 = (int(v) for v in g[f"{tag}_params"])
    out = oracle.correlation(T(g[f"{tag}_f1"]), T(g[f"{tag}_f2"]), pad, k, md, s1, s2)
    assert out.shape == g[f"{tag}_out"].shape
    np.testing.assert_allclose(out.numpy(), g[f"{tag}_out"], atol=1e-14, rtol=0)


def test_correlation_loops_agree():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(1, 3, 5, 6, generator=g, dtype=torch.float64)
    b = torch.randn(1, 3, 5, 6, generator=g, dtype=torch.float64)
    for pad, k, md, s1, s2 in [(2, 1, 2, 1, 1), (4, 1, 4, 1, 2), (3, 3, 2, 1, 1), (2, 1, 2, 2, 1)]:
        ref = oracle.correlation_loops(a.numpy(), b.numpy(), pad, k, md, s1, s2)
        out = oracle.correlation(a, b, pad, k, md, s1, s2)
        np.testing.assert_allclose(out.numpy(), ref, atol=1e-14)


def test_correlation_one_hot_picks_shift():
    # one-hot feature at (y0,x0) in f1 and at (y0+2*tj, x0+2*ti) in f2 -> only channel (tj+r)*D+(ti+r) fires
    H = W = 12
    md, s2 = 4, 2
    r, D = md // s2, 2 * (md // s2) + 1
    f1 = torch.zeros(1, 1, H, W, dtype=torch.float64)
    f2 = torch.zeros(1, 1, H, W, dtype=torch.float64)
    y0, x0, tj, ti = 5, 6, -1, 2
    f1[0, 0, y0, x0] = 1.0
    f2[0, 0, y0 + s2 * tj, x0 + s2 * ti] = 1.0
    out = oracle.correlation(f1, f2, md, 1, md, 1, s2)
    assert out.shape == (1, D * D, H, W)
    nz = out.nonzero()
    assert nz.shape[0] == 1
    assert tuple(nz[0].tolist()) == (0, (tj + r) * D + (ti + r), y0, x0)


def test_correlation_shapes_reference_configs():
    assert oracle.correlation_out_shape(32, 32, 20, 1, 20, 1, 2) == (21, 32, 32)  # FlowNetC.py:31
    assert oracle.correlation_out_shape(8, 8, 4, 1, 4, 1, 1) == (9, 8, 8)          # PWCNet.py:69


# ----------------------------------------------------------------------------- resample2d
@pytest.mark.parametrize("tag", ["a", "b"])
def test_resample2d_vs_grid_sample_golden(golden, tag):
    g = golden("resample")
    out = oracle.resample2d(T(g[f"{tag}_in1"]), T(g[f"{tag}_flow"]))
    np.testing.assert_allclose(out.numpy(), g[f"{tag}_out"], atol=2e-12)
    loops = oracle.resample2d_loops(g[f"{tag}_in1"], g[f"{tag}_flow"])
    np.testing.assert_allclose(out.numpy(), loops, atol=1e-13)


def test_resample2d_identity_and_integer_shift():
    g = torch.Generator().manual_seed(2)
    img = torch.randn(2, 3, 9, 11, generator=g, dtype=torch.float64)
    z = torch.zeros(2, 2, 9, 11, dtype=torch.float64)
    assert torch.equal(oracle.resample2d(img, z), img)
    fl = z.clone()
    fl[:, 0] = 2.0
    fl[:, 1] = -1.0
    out = oracle.resample2d(img, fl)
    assert torch.equal(out[:, :, 1:, :-2], img[:, :, :-1, 2:])
    # border clamp: columns beyond the image repeat the last column
    assert torch.equal(out[:, :, 1:, -1], img[:, :, :-1, -1])


def test_resample2d_nearest():
    img = torch.arange(16, dtype=torch.float64).view(1, 1, 4, 4)
    fl = torch.zeros(1, 2, 4, 4, dtype=torch.float64)
    fl[:, 0] = 0.6
    out = oracle.resample2d(img, fl, bilinear=False)
    assert torch.equal(out[0, 0, :, :3], img[0, 0, :, 1:])


# ----------------------------------------------------------------------------- channelnorm
def test_channelnorm():
    x = torch.tensor([[[[3.0]], [[4.0]]]], dtype=torch.float64)
    out = oracle.channelnorm(x)
    assert out.item() == 5.0
    g = oracle.channelnorm_backward(x, out, torch.ones_like(out))
    np.testing.assert_allclose(g.flatten().numpy(), [0.6, 0.8], rtol=1e-8)
    z = torch.zeros(1, 2, 2, 2, dtype=torch.float64)
    assert torch.equal(oracle.channelnorm_backward(z, oracle.channelnorm(z), torch.ones(1, 1, 2, 2).double()), z)


# ----------------------------------------------------------------------------- warps (reference-pinned)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stn_vs_reference_golden(golden, tag):
    g = golden("warps")
    out = oracle.stn_warp(T(g[f"stn_flow_{tag}"], torch.float32), T(g["stn_frame"], torch.float32))
    np.testing.assert_allclose(out.numpy(), g[f"stn_out_{tag}"], atol=1e-6)


def test_stn_zero_flow_is_not_identity(golden):
    g = golden("warps")
    frame = T(g["stn_frame"], torch.float32)[:1, :, :8, :8].contiguous()
    out = oracle.stn_warp(torch.zeros(1, 2, 8, 8), frame)
    np.testing.assert_allclose(out.numpy(), g["stn_zero"], atol=1e-6)
    # samples at x*(W-1)/W
    xs = torch.arange(8, dtype=torch.float64) * 7 / 8
    x0 = xs.floor().long()
    a = xs - x0
    row = frame[0, 0].double()
    x1 = (x0 + 1).clamp(max=7)
    interp_x = row[:, x0] * (1 - a) + row[:, x1] * a
    interp = interp_x[x0] * (1 - a).view(-1, 1) + interp_x[x1] * a.view(-1, 1)
    np.testing.assert_allclose(out[0, 0].double().numpy(), interp.numpy(), atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pwc_warp_vs_reference_golden(golden, tag):
    g = golden("warps")
    out = oracle.pwc_warp(T(g[f"pwc_x_{tag}"], torch.float32), T(g[f"pwc_flow_{tag}"], torch.float32))
    np.testing.assert_allclose(out.numpy(), g[f"pwc_out_{tag}"], atol=1e-6)


def test_pwc_zero_flow_masks_border(golden):
    g = golden("warps")
    out = oracle.pwc_warp(T(g["pwc_zero_x"], torch.float32), torch.zeros(1, 2, 8, 8))
    np.testing.assert_allclose(out.numpy(), g["pwc_zero_out"], atol=1e-6)
    assert int((out == 0).sum()) == 28  # the whole 1-px border


# ----------------------------------------------------------------------------- loss (reference-pinned)
def test_loss_terms_vs_reference_golden(golden):
    g = golden("loss")
    fixed = T(g["fixed"], torch.float32)
    for i in range(4):
        w = T(g[f"warped{i}"], torch.float32)
        f = T(g[f"flow{i}"], torch.float32)
        np.testing.assert_allclose(oracle.photometric(fixed, w).numpy(), g[f"photo{i}"], rtol=1e-6)
        np.testing.assert_allclose(oracle.ncc(fixed, w).numpy(), g[f"ncc{i}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(oracle.smoothness(f).numpy(), g[f"smooth{i}"], rtol=1e-6)


def test_ofe_loss_vs_reference_golden(golden):
    g = golden("loss")
    fixed = T(g["fixed"], torch.float32)
    flows = [T(g[f"flow{i}"], torch.float32) for i in range(4)]
    warped = [T(g[f"warped{i}"], torch.float32) for i in range(4)]
    res = oracle.ofe_loss(flows, warped, fixed)
    assert all(t.dtype == torch.float64 for t in res)
    np.testing.assert_allclose([t.item() for t in res], g["ofe"], rtol=1e-6)
    res = oracle.ofe_loss(flows[:2], warped[:2], fixed, lamb_da=0.25, gamma=10.0, zeta=3.0)
    np.testing.assert_allclose([t.item() for t in res], g["ofe_alt"], rtol=1e-6)


def test_loss_gradients_vs_reference_golden(golden):
    g = golden("loss")
    fixed = T(g["fixed"], torch.float32)
    flows = [T(g[f"flow{i}"], torch.float32).requires_grad_(True) for i in range(4)]
    warped = [T(g[f"warped{i}"], torch.float32).requires_grad_(True) for i in range(4)]
    oracle.ofe_loss(flows, warped, fixed)[3].backward()
    for i in range(4):
        np.testing.assert_allclose(flows[i].grad.numpy(), g[f"gflow{i}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(warped[i].grad.numpy(), g[f"gwarped{i}"], rtol=1e-5, atol=1e-6)


def test_registration_chain_vs_reference_golden(golden):
    """The reference's own chain stn -> OFEloss (models.py:277, train.py:56), values and d total / d flow through both
    (regloss.npz): the oracle restatement in float32, the reference's arithmetic."""
    g = golden("regloss")
    fixed, moving = T(g["fixed"], torch.float32), T(g["moving"], torch.float32)
    flows = [T(g[f"flow{i}"], torch.float32).requires_grad_(True) for i in range(4)]
    warped = [oracle.stn_warp(f, moving) for f in flows]
    res = oracle.ofe_loss(flows, warped, fixed)
    np.testing.assert_allclose([float(t.detach()) for t in res], g["ofe"], rtol=2e-6)
    res[3].backward()
    for i in range(4):
        np.testing.assert_allclose(warped[i].detach().numpy(), g[f"warped{i}"], atol=1e-6)
        np.testing.assert_allclose(flows[i].grad.numpy(), g[f"gflow{i}"], rtol=1e-4, atol=1e-5)


def test_loss_degenerate_cases(golden):
    g = golden("loss")
    fixed = T(g["fixed"], torch.float32)
    const = torch.full((3, 1, 16, 16), 0.5)
    assert float(oracle.ncc(fixed, const)) == float(g["ncc_const"]) == 0.0
    z = oracle.smoothness(torch.zeros(3, 2, 8, 8))
    np.testing.assert_allclose(z.numpy(), g["smooth_zero"], rtol=1e-6)
    # floor value: 2 channels * 64 px * (1e-18)^0.25, /2 for the pair... per image
    np.testing.assert_allclose(float(z), 2 * 64 * (1e-18) ** 0.25, rtol=1e-5)


def test_ncc_constant_images_vs_reference_golden(golden):
    """loss.py:58-60 on constant warped images (ncc_const.npz, produced by the reference): loss 0 when the float32 mean
    is exact, ~1 otherwise (0.7).  The oracle reproduces both; the CUDA kernels return 0 for every constant image
    (documented divergence, include/ofr_b200.h) -- tests/test_gpu_parity.py::test_ncc_constant_image_rule."""
    g = golden("ncc_const")
    fixed = torch.from_numpy(g["fixed"])
    for v, want, zero in zip(g["values"], g["ref_loss"], g["centred_is_zero"]):
        got = float(oracle.ncc(fixed, torch.full((3, 1, 16, 16), float(v))))
        assert abs(got - want) <= 1e-6, (v, got, want)
        assert (want == 0.0) == bool(zero)
    assert g["ref_loss"][3] > 0.99   # 0.7: float32 rounding noise decides in the reference


@pytest.mark.parametrize("shape", [(2, 5, 6, 8), (1, 3, 1, 4), (2, 2, 7, 5)])
def test_flow_head_loop_restatements_equal_torch_convolutions(shape):
    """the reference's flow heads are ``nn.Conv2d(C, 2, 3, 1, 1)`` / ``nn.ConvTranspose2d(C, 2, 4, 2, 1)``
    (PWCNet.py:30-34); the explicit shifted-sum statements in oracle/flowhead_ref.py agree with PyTorch's ops in float64."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w3 = torch.randn(2, C, 3, 3, generator=g, dtype=torch.float64)
    w4 = torch.randn(C, 2, 4, 4, generator=g, dtype=torch.float64)
    b = torch.randn(2, generator=g, dtype=torch.float64)
    assert (oracle.predict_flow_loops(x, w3, b) - oracle.predict_flow(x, w3, b)).abs().max().item() < 1e-12
    assert (oracle.deconv_loops(x, w4, b) - oracle.deconv(x, w4, b)).abs().max().item() < 1e-12
    assert oracle.deconv(x, w4).shape == (B, 2, 2 * H, 2 * W)


@pytest.mark.parametrize("kind,tag", [("conv", "a"), ("conv", "b"), ("deconv", "a"), ("deconv", "b")])
def test_flow_head_oracle_vs_reference_golden(golden, kind, tag):
    """flowhead.npz holds what the reference's own ``predict_flow`` / ``deconv`` layers (PWCNet.py:30-34) computed on the
    CPU in float32 (tests/golden/make_golden.py --only-flowhead); the oracle, run in float64, agrees to float32 rounding."""
    g = golden("flowhead")
    k = f"{kind}_{tag}_"
    x, w, b, go = (torch.from_numpy(g[k + n]).double() for n in ("x", "w", "b", "go"))
    x.requires_grad_(True), w.requires_grad_(True), b.requires_grad_(True)
    y = (oracle.predict_flow if kind == "conv" else oracle.deconv)(x, w, b)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), go)
    for name, got, ref in (("y", y, g[k + "y"]), ("gx", gx, g[k + "gx"]), ("gw", gw, g[k + "gw"]), ("gb", gb, g[k + "gb"])):
        ref = torch.from_numpy(ref).double()
        assert (got.detach() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), name

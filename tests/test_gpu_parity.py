"""GPU parity: every kernel, through the C ABI (ctypes -> libofr_b200.so), against the CPU oracle.

Tolerances are the north-star's: max abs error <= 1e-5 on correlation and warp outputs, <= 1e-4
relative (to the gradient's max magnitude) on gradients.  The oracle runs in float64 on the CPU.
"""
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
    ofr_b200._lib.lib()  # must load: no fallback
    return ofr_b200


def gen(seed):
    return torch.Generator().manual_seed(seed)


def close(got, ref, atol=0.0, rtol=0.0, what=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item() if ref.numel() else 0.0
    bound = atol + rtol * (ref.abs().max().item() if ref.numel() else 0.0)
    assert err <= bound, f"{what}: max abs err {err:.3e} > bound {bound:.3e}"


def run_pair(ref_fn, dev_fn, inputs, out_atol, grad_rtol, what, grad_atol=1e-7):
    """Forward + backward of ref_fn on CPU/float64 and dev_fn on the GPU/float32 with a shared cotangent."""
    cpu = [t.detach().double().requires_grad_(True) for t in inputs]
    ref = ref_fn(*cpu)
    go = torch.randn(ref.shape, generator=gen(99), dtype=torch.float64)
    rg = torch.autograd.grad(ref, cpu, go, allow_unused=True)
    dev = [t.detach().float().to(DEV).requires_grad_(True) for t in inputs]
    out = dev_fn(*dev)
    dg = torch.autograd.grad(out, dev, go.float().to(DEV), allow_unused=True)
    close(out, ref, atol=out_atol, what=what + " fwd")
    for i, (g, r) in enumerate(zip(dg, rg)):
        if r is None:
            continue
        close(g, r, atol=grad_atol, rtol=grad_rtol, what=f"{what} grad[{i}]")
    return out, dg


# =============================================================================== ChannelNorm
@pytest.mark.parametrize("shape", [(2, 1, 16, 16), (3, 2, 8, 12), (2, 3, 7, 9), (1, 5, 1, 1), (2, 2, 33, 5)])
def test_channelnorm(ofr, shape):
    x = torch.randn(*shape, generator=gen(1))
    xd = x.to(DEV).requires_grad_(True)
    out = ofr.ChannelNorm()(xd)
    close(out, oracle.channelnorm(x.double()), atol=1e-6, what="chnorm fwd")
    go = torch.randn(out.shape, generator=gen(2))
    (g,) = torch.autograd.grad(out, xd, go.to(DEV))
    ref = oracle.channelnorm_backward(x.double(), oracle.channelnorm(x.double()), go.double())
    close(g, ref, atol=1e-7, rtol=1e-5, what="chnorm bwd")


def test_channelnorm_zero_input_has_finite_gradient(ofr):
    xd = torch.zeros(1, 2, 4, 4, device=DEV, requires_grad=True)
    out = ofr.ChannelNorm()(xd)
    (g,) = torch.autograd.grad(out.sum(), xd)
    assert torch.equal(out, torch.zeros_like(out)) and torch.equal(g, torch.zeros_like(g))


# =============================================================================== Correlation
CORR_CASES = {
    # tag: (B, C, H, W, pad, k, md, s1, s2)
    "pwc_l6": (2, 196, 4, 4, 4, 1, 4, 1, 1),
    "pwc_l5": (2, 128, 8, 8, 4, 1, 4, 1, 1),
    "pwc_l4": (2, 96, 16, 16, 4, 1, 4, 1, 1),
    "pwc_l3": (1, 64, 32, 32, 4, 1, 4, 1, 1),
    "pwc_l2": (1, 32, 64, 64, 4, 1, 4, 1, 1),
    "pwc_ragged": (2, 13, 12, 20, 4, 1, 4, 1, 1),      # C not a multiple of the chunk, H not of the tile
    "pwc_wide": (1, 8, 6, 136, 4, 1, 4, 1, 1),          # several x tiles with a ragged last one
    "pwc_ragged_big": (3, 40, 50, 72, 4, 1, 4, 1, 1),    # 2 channel blocks (8 ragged), ragged last row tile and column strip
    "pwc_tall": (2, 8, 200, 32, 4, 1, 4, 1, 1),           # 25 row tiles of one column strip
    "flownetc": (1, 256, 32, 32, 20, 1, 20, 1, 2),
    "flownetc_small": (2, 20, 12, 16, 20, 1, 20, 1, 2),
    "flownetc_tc": (2, 48, 32, 64, 20, 1, 20, 1, 2),    # class images 16 x 32: tensor-core forward, 2 x 2 tiles, 3 K stages
    "generic_w_odd": (2, 5, 9, 11, 4, 1, 4, 1, 1),      # W % 4 != 0 -> generic kernels
    "generic_k3": (1, 3, 9, 10, 3, 3, 2, 1, 1),
    "generic_s1": (1, 4, 12, 12, 4, 1, 4, 2, 2),
    "generic_pad_lt_md": (1, 4, 16, 16, 2, 1, 4, 1, 1),  # output smaller than input
    "generic_pad_gt_md": (1, 3, 8, 8, 6, 1, 4, 1, 2),    # output larger than input
    "generic_s2_3": (1, 6, 12, 12, 6, 1, 6, 1, 3),
}


@pytest.mark.parametrize("tag", list(CORR_CASES))
def test_correlation_fwd_bwd(ofr, tag):
    B, C, H, W, pad, k, md, s1, s2 = CORR_CASES[tag]
    f1 = torch.randn(B, C, H, W, generator=gen(3))
    f2 = torch.randn(B, C, H, W, generator=gen(4))
    run_pair(lambda a, b: oracle.correlation(a, b, pad, k, md, s1, s2),
             lambda a, b: ofr.Correlation(pad, k, md, s1, s2, 1)(a, b),
             [f1, f2], out_atol=1e-5, grad_rtol=1e-4, what=f"corr[{tag}]")


@pytest.mark.parametrize("tag", ["pwc", "fnc", "k3", "s1"])
def test_correlation_golden(ofr, golden, tag):
    g = golden("corr")
    pad, k, md, s1, s2 = (int(v) for v in g[f"{tag}_params"])
    out = ofr.Correlation(pad, k, md, s1, s2, 1)(torch.from_numpy(g[f"{tag}_f1"]).to(DEV),
                                                 torch.from_numpy(g[f"{tag}_f2"]).to(DEV))
    close(out, torch.from_numpy(g[f"{tag}_out"]), atol=1e-5, what=f"corr golden {tag}")


def test_correlation_one_hot(ofr):
    H = W = 16
    for md, s2 in [(4, 1), (20, 2)]:
        r, D = md // s2, 2 * (md // s2) + 1
        f1 = torch.zeros(1, 4, H, W)
        f2 = torch.zeros(1, 4, H, W)
        y0, x0, tj, ti = 7, 8, -2, 3
        f1[0, 2, y0, x0] = 2.0
        f2[0, 2, y0 + s2 * tj, x0 + s2 * ti] = 3.0
        out = ofr.Correlation(md, 1, md, 1, s2, 1)(f1.to(DEV), f2.to(DEV)).cpu()
        nz = out.nonzero()
        assert nz.shape[0] == 1 and tuple(nz[0].tolist()) == (0, (tj + r) * D + (ti + r), y0, x0)
        assert abs(out[tuple(nz[0].tolist())].item() - 6.0 / 4) < 1e-7


def test_correlation_flownetc_full_size_tensor_core_forward(ofr):
    """Correlation(20,1,20,1,2) at the FlowNetC shape (B=24, C=256, 64x64: FlowNetC.py:31,88) -- the tcgen05 3xTF32 path
    -- against the definition evaluated in float64 on the GPU for a sample of displacement planes (centre, corners,
    mixed), with the fused LeakyReLU; exact on one-hot features; bitwise reproducible."""
    B, C, H, W, md, s2 = 24, 256, 64, 64, 20, 2
    g0 = torch.Generator(device=DEV).manual_seed(90)
    f1 = torch.randn(B, C, H, W, device=DEV, generator=g0)
    f2 = torch.randn(B, C, H, W, device=DEV, generator=g0)
    corr = ofr.Correlation(md, 1, md, 1, s2, 1, leaky_relu=0.1)
    out = corr(f1, f2)
    assert torch.equal(out, corr(f1, f2))
    r, D = md // s2, 2 * (md // s2) + 1
    f1d, f2d = f1.double(), torch.nn.functional.pad(f2.double(), (md, md, md, md))
    for tj, ti in [(0, 0), (-r, -r), (r, r), (-r, r), (3, -7), (-1, 1), (r, 0), (0, -r)]:
        sh = f2d[:, :, md + s2 * tj: md + s2 * tj + H, md + s2 * ti: md + s2 * ti + W]
        ref = torch.nn.functional.leaky_relu((f1d * sh).mean(1), 0.1)
        close(out[:, (tj + r) * D + (ti + r)], ref, atol=1e-5, what=f"flownetc full-size plane ({tj},{ti})")
    a = torch.zeros(1, 16, 32, 32, device=DEV)
    b = torch.zeros(1, 16, 32, 32, device=DEV)
    a[0, 5, 9, 20] = 2.0
    b[0, 5, 9 + s2 * 4, 20 - s2 * 9] = 3.0
    o = ofr.Correlation(md, 1, md, 1, s2, 1)(a, b)
    nz = o.nonzero()
    assert nz.shape[0] == 1 and tuple(nz[0].tolist()) == (0, (4 + r) * D + (-9 + r), 9, 20)
    assert o[tuple(nz[0].tolist())].item() == 6.0 / 16


def test_correlation_flownetc_full_size_tensor_core_backward(ofr):
    """Backward of Correlation(20,1,20,1,2) at the FlowNetC shape (B=24, C=256, 64x64) -- ``corr_bwd_umma_kernel``:
    in-place stride-2 addressing, row skipping at the top/bottom tiles, 256 TMEM columns.  (i) Euler identity of the
    bilinear operator, <dL/df1, f1> == <gO, out> == <dL/df2, f2>; (ii) a sample of channels of both gradients against
    the definition accumulated over all 441 displacements in float64 on the GPU; (iii) bitwise reproducible."""
    B, C, H, W, md, s2 = 24, 256, 64, 64, 20, 2
    g0 = torch.Generator(device=DEV).manual_seed(91)
    f1 = torch.randn(B, C, H, W, device=DEV, generator=g0, requires_grad=True)
    f2 = torch.randn(B, C, H, W, device=DEV, generator=g0, requires_grad=True)
    corr = ofr.Correlation(md, 1, md, 1, s2, 1)
    out = corr(f1, f2)
    go = torch.randn(out.shape, device=DEV, generator=g0)
    g1, g2 = torch.autograd.grad(out, (f1, f2), go, retain_graph=True)
    h1, h2 = torch.autograd.grad(out, (f1, f2), go)
    assert torch.equal(g1, h1) and torch.equal(g2, h2)
    lhs = (go.double() * out.detach().double()).sum().item()
    for name, g, f in (("f1", g1, f1), ("f2", g2, f2)):
        rhs = (g.double() * f.detach().double()).sum().item()
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (name, lhs, rhs)
    r, D = md // s2, 2 * (md // s2) + 1
    cs = [0, 7, 100, 255]
    a = f1.detach()[:, cs].double()
    b = torch.nn.functional.pad(f2.detach()[:, cs].double(), (md, md, md, md))
    ref1 = torch.zeros(B, len(cs), H, W, dtype=torch.float64, device=DEV)
    ref2 = torch.zeros(B, len(cs), H + 2 * md, W + 2 * md, dtype=torch.float64, device=DEV)
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            gd = go[:, (tj + r) * D + (ti + r)].double().unsqueeze(1)
            ys, xs = md + s2 * tj, md + s2 * ti
            ref1 += gd * b[:, :, ys:ys + H, xs:xs + W]
            ref2[:, :, ys:ys + H, xs:xs + W] += gd * a
    ref1 /= C
    ref2 = ref2[:, :, md:md + H, md:md + W] / C
    close(g1[:, cs], ref1, atol=1e-7, rtol=1e-4, what="flownetc full-size dI1")
    close(g2[:, cs], ref2, atol=1e-7, rtol=1e-4, what="flownetc full-size dI2")


def test_resample2d_full_size(ofr):
    """Resample2d (mode 0: pixel-unit flow, border-clamped taps; flownet2/models.py:139-176) at 24 x 1 x 512 x 512:
    forward and both gradients against the oracle in the layer's own float32 arithmetic, bitwise reproducible."""
    B, S = 24, 512
    g0 = torch.Generator(device=DEV).manual_seed(92)
    img = torch.rand(B, 1, S, S, device=DEV, generator=g0)
    coarse = torch.randn(B, 2, S // 16, S // 16, device=DEV, generator=g0) * 6
    flow = torch.nn.functional.interpolate(coarse, size=(S, S), mode="bilinear", align_corners=False)
    go = torch.randn(B, 1, S, S, device=DEV, generator=g0)
    xi, fi = img.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    out = ofr.Resample2d()(xi, fi)
    gx, gf = torch.autograd.grad(out, (xi, fi), go, retain_graph=True)
    hx, hf = torch.autograd.grad(out, (xi, fi), go)
    assert torch.equal(gx, hx) and torch.equal(gf, hf)
    xc, fc = img.cpu().requires_grad_(True), flow.cpu().requires_grad_(True)
    ref = oracle.resample2d(xc, fc)
    rgx, rgf = torch.autograd.grad(ref, (xc, fc), go.cpu())
    close(out, ref, atol=1e-5, what="full-size resample2d")
    close(gx, rgx, atol=1e-6, rtol=1e-4, what="full-size resample2d dImg")
    close(gf, rgf, atol=1e-5, rtol=1e-4, what="full-size resample2d dFlow")
    lhs = (go.double() * out.detach().double()).sum().item()
    rhs = (gx.double() * img.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)


def test_correlation_is_reproducible_and_skips_unneeded_grads(ofr):
    f1 = torch.randn(2, 32, 16, 16, generator=gen(5)).to(DEV).requires_grad_(True)
    f2 = torch.randn(2, 32, 16, 16, generator=gen(6)).to(DEV)
    corr = ofr.Correlation(4, 1, 4, 1, 1, 1)
    o1, o2 = corr(f1, f2), corr(f1, f2)
    assert torch.equal(o1, o2)
    (g,) = torch.autograd.grad(o1.sum(), f1)
    (g2,) = torch.autograd.grad(o2.sum(), f1)
    assert torch.equal(g, g2)


@pytest.mark.parametrize("which", [0, 1])
def test_correlation_backward_single_gradient_many_tiles(ofr, which):
    """PWC backward with only ONE gradient requested (the launch then holds one direction) on a batch with far more tiles
    than SMs (B = 40, 32 channels, 48 x 64): against the definition accumulated in float64 on the GPU, and bit-identical
    run to run."""
    B, C, H, W = 40, 32, 48, 64
    f1 = torch.randn(B, C, H, W, generator=gen(130)).to(DEV)
    f2 = torch.randn(B, C, H, W, generator=gen(131)).to(DEV)
    go = torch.randn(B, 81, H, W, generator=gen(132)).to(DEV)
    xs = [f1.clone().requires_grad_(which == 0), f2.clone().requires_grad_(which == 1)]
    corr = ofr.Correlation(4, 1, 4, 1, 1, 1)
    (g,) = torch.autograd.grad(corr(*xs), xs[which], go)
    (g2,) = torch.autograd.grad(corr(*xs), xs[which], go)
    assert torch.equal(g, g2)
    # float64 definition on the GPU
    pad = torch.nn.functional.pad
    a, b, gd = f1.double(), f2.double(), go.double()
    ref = torch.zeros(B, C, H + 8, W + 8, dtype=torch.float64, device=DEV) if which == 1 else torch.zeros(B, C, H, W, dtype=torch.float64, device=DEV)
    bp = pad(b, (4, 4, 4, 4))
    for tj in range(9):
        for ti in range(9):
            gp = gd[:, tj * 9 + ti].unsqueeze(1)
            if which == 0:
                ref += gp * bp[:, :, tj:tj + H, ti:ti + W]
            else:
                ref[:, :, tj:tj + H, ti:ti + W] += gp * a
    ref = (ref if which == 0 else ref[:, :, 4:4 + H, 4:4 + W]) / C
    close(g, ref, atol=1e-7, rtol=1e-4, what=f"single-gradient backward grad[{which}]")


def test_spatial_correlation_sample_alias(ofr):
    # FlowNetS/util.py:58-72: patch 21, dilation_patch 2, then /C  ==  Correlation(20,1,20,1,2,1)
    a = torch.randn(1, 6, 12, 12, generator=gen(7)).to(DEV)
    b = torch.randn(1, 6, 12, 12, generator=gen(8)).to(DEV)
    s = ofr.spatial_correlation_sample(a, b, kernel_size=1, patch_size=21, stride=1, padding=0, dilation_patch=2)
    assert s.shape == (1, 21, 21, 12, 12)
    close(s.view(1, 441, 12, 12) / 6, ofr.Correlation(20, 1, 20, 1, 2, 1)(a, b), atol=1e-6, what="scs alias")


# =============================================================================== warps
WARP_SHAPES = [(2, 3, 16, 16), (1, 1, 9, 13), (2, 17, 8, 12), (1, 5, 32, 20)]


@pytest.mark.parametrize("shape", WARP_SHAPES)
@pytest.mark.parametrize("mode", ["resample2d", "stn", "pwc"])
def test_warp_fwd_bwd(ofr, mode, shape):
    B, C, H, W = shape
    img = torch.randn(B, C, H, W, generator=gen(10))
    flow = torch.randn(B, 2, H, W, generator=gen(11)) * 3.0
    fns = {
        "resample2d": (lambda i, f: oracle.resample2d(i, f), lambda i, f: ofr.Resample2d()(i, f)),
        "stn": (lambda i, f: oracle.stn_warp(f, i), lambda i, f: ofr.stn_warp(f, i)),
        "pwc": (lambda i, f: oracle.pwc_warp(i, f), lambda i, f: ofr.pwc_warp(i, f)),
    }
    run_pair(fns[mode][0], fns[mode][1], [img, flow], out_atol=1e-5, grad_rtol=1e-4, what=f"{mode}{shape}",
             grad_atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_resample2d_golden(ofr, golden, tag):
    g = golden("resample")
    out = ofr.Resample2d()(torch.from_numpy(g[f"{tag}_in1"]).to(DEV), torch.from_numpy(g[f"{tag}_flow"]).to(DEV))
    close(out, torch.from_numpy(g[f"{tag}_out"]), atol=1e-5, what="resample golden")


def test_warps_vs_reference_golden(ofr, golden):
    g = golden("warps")
    frame = torch.from_numpy(g["stn_frame"]).to(DEV)
    for tag in "abc":
        out = ofr.stn_warp(torch.from_numpy(g[f"stn_flow_{tag}"]).to(DEV), frame)
        close(out, torch.from_numpy(g[f"stn_out_{tag}"]), atol=1e-5, what=f"stn golden {tag}")
    for tag in "ab":
        out = ofr.pwc_warp(torch.from_numpy(g[f"pwc_x_{tag}"]).to(DEV), torch.from_numpy(g[f"pwc_flow_{tag}"]).to(DEV))
        close(out, torch.from_numpy(g[f"pwc_out_{tag}"]), atol=1e-5, what=f"pwc golden {tag}")
    out = ofr.pwc_warp(torch.from_numpy(g["pwc_zero_x"]).to(DEV), torch.zeros(1, 2, 8, 8, device=DEV))
    close(out, torch.from_numpy(g["pwc_zero_out"]), atol=1e-6, what="pwc zero flow")
    assert int((out == 0).sum()) == 28


def test_resample2d_identity_shift_and_sizes(ofr):
    img = torch.randn(2, 3, 12, 16, generator=gen(12)).to(DEV)
    z = torch.zeros(2, 2, 12, 16, device=DEV)
    assert torch.equal(ofr.Resample2d()(img, z), img)
    fl = z.clone()
    fl[:, 0] = 3.0
    fl[:, 1] = -2.0
    out = ofr.Resample2d()(img, fl)
    assert torch.equal(out[:, :, 2:, :-3], img[:, :, :-2, 3:])
    assert torch.equal(out[:, :, 2:, -1], img[:, :, :-2, -1])  # clamped border
    # flow smaller than the image: output takes the flow's size, taps clamp to the image
    fl2 = torch.randn(2, 2, 6, 8, generator=gen(13)) * 4
    ref = oracle.resample2d(img.cpu().double(), fl2.double())
    close(ofr.Resample2d()(img, fl2.to(DEV)), ref, atol=1e-5, what="resample Hin != Hf")
    # nearest mode
    refn = oracle.resample2d(img.cpu().double(), fl2.double(), bilinear=False)
    close(ofr.Resample2d(bilinear=False)(img, fl2.to(DEV)), refn, atol=0, what="resample nearest")


@pytest.mark.parametrize("mode", ["resample2d", "pwc", "stn"])
def test_warp_backward_is_bitwise_deterministic(ofr, mode):
    """North-star rule 3: the adjoint scatter must not depend on execution order."""
    B, C, H, W = 2, 24, 32, 32
    img = torch.randn(B, C, H, W, generator=gen(14)).to(DEV).requires_grad_(True)
    flow = (torch.randn(B, 2, H, W, generator=gen(15)) * 4).to(DEV).requires_grad_(True)
    go = torch.randn(B, C, H, W, generator=gen(16)).to(DEV)
    fn = {"resample2d": lambda: ofr.Resample2d()(img, flow), "pwc": lambda: ofr.pwc_warp(img, flow),
          "stn": lambda: ofr.stn_warp(flow, img)}[mode]
    first = torch.autograd.grad(fn(), (img, flow), go)
    for _ in range(4):
        again = torch.autograd.grad(fn(), (img, flow), go)
        assert torch.equal(first[0], again[0]) and torch.equal(first[1], again[1])


@pytest.mark.parametrize("kind", ["collapse_point", "collapse_row", "far_outside", "constant_shift"])
def test_warp_backward_degenerate_flows(ofr, kind):
    """Crowded cells (many sources on one target) take the pre-sorted path; results stay exact and reproducible."""
    B, C, H, W = 1, 9, 24, 40
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    flow = torch.zeros(B, 2, H, W)
    if kind == "collapse_point":      # every pixel samples around (5.3, 7.6)
        flow[:, 0], flow[:, 1] = 5.3 - xs, 7.6 - ys
    elif kind == "collapse_row":      # every pixel of a row samples column 3.5
        flow[:, 0] = 3.5 - xs
    elif kind == "far_outside":       # everything clamps to the bottom-right corner / falls outside
        flow[:] = 1000.0
    else:
        flow[:, 0], flow[:, 1] = 2.25, -1.5
    img = torch.randn(B, C, H, W, generator=gen(17))
    for ref_fn, dev_fn, name in [(lambda i, f: oracle.resample2d(i, f), lambda i, f: ofr.Resample2d()(i, f), "resample2d"),
                                 (lambda i, f: oracle.pwc_warp(i, f), lambda i, f: ofr.pwc_warp(i, f), "pwc")]:
        cpu = [img.double().requires_grad_(True), flow.double().requires_grad_(True)]
        ref = ref_fn(*cpu)
        go = torch.randn(ref.shape, generator=gen(18), dtype=torch.float64)
        (rg,) = torch.autograd.grad(ref, cpu[:1], go)
        dev = [img.to(DEV).requires_grad_(True), flow.to(DEV)]
        g1 = torch.autograd.grad(dev_fn(*dev), dev[:1], go.float().to(DEV))[0]
        g2 = torch.autograd.grad(dev_fn(*dev), dev[:1], go.float().to(DEV))[0]
        assert torch.equal(g1, g2)
        close(g1, rg, atol=1e-5, rtol=2e-4, what=f"{name} {kind} dImg")


# =============================================================================== fused epilogues
@pytest.mark.parametrize("tag", ["pwc_l6", "pwc_l4", "pwc_l2", "flownetc_small", "generic_w_odd"])
def test_correlation_fused_leaky_relu(ofr, tag):
    """Correlation(..., leaky_relu=0.1) == LeakyReLU(0.1)(Correlation(...)) (PWCNet.py:200-201), values and gradients."""
    B, C, H, W, pad, k, md, s1, s2 = CORR_CASES[tag]
    a = torch.randn(B, C, H, W, generator=gen(30)).to(DEV)
    b = torch.randn(B, C, H, W, generator=gen(31)).to(DEV)
    outs, grads = [], []
    for fused in (False, True):
        x, y = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        layer = ofr.Correlation(pad, k, md, s1, s2, 1, leaky_relu=0.1 if fused else None)
        out = layer(x, y)
        if not fused:
            out = torch.nn.functional.leaky_relu(out, 0.1)
        go = torch.randn(out.shape, generator=gen(32)).to(DEV)
        outs.append(out.detach())
        grads.append(torch.autograd.grad(out, (x, y), go))
    assert torch.equal(outs[0], outs[1])
    assert (outs[0] < 0).any()
    for g0, g1 in zip(grads[0], grads[1]):
        assert torch.equal(g0, g1)


def test_pwc_warp_fused_flow_scale(ofr):
    """pwc_warp(x, flo, s) == pwc_warp(x, flo * s) (PWCNet.py:214 passes `up_flow * 0.625`), values and gradients."""
    B, C, H, W = 2, 24, 32, 32
    x = torch.randn(B, C, H, W, generator=gen(33)).to(DEV)
    fl = (torch.randn(B, 2, H, W, generator=gen(34)) * 2).to(DEV)
    go = torch.randn(B, C, H, W, generator=gen(35)).to(DEV)
    res = []
    for fused in (False, True):
        xi, fi = x.clone().requires_grad_(True), fl.clone().requires_grad_(True)
        out = ofr.pwc_warp(xi, fi, 2.5) if fused else ofr.pwc_warp(xi, fi * 2.5)
        res.append((out.detach(),) + torch.autograd.grad(out, (xi, fi), go))
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
    close(res[1][2], res[0][2], atol=1e-6, rtol=1e-6, what="flow gradient through the fused scale")


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 1, 7, 9), (2, 2, 33, 5)])
def test_channelnorm_diff_equals_channelnorm_of_the_difference(ofr, shape):
    """channelnorm_diff(a, b) == ChannelNorm()(a - b) (flownet2/models.py:140-141), values and both gradients."""
    a0 = torch.randn(*shape, generator=gen(90)).to(DEV)
    b0 = torch.randn(*shape, generator=gen(91)).to(DEV)
    b0[0, :, 0, 0] = a0[0, :, 0, 0]  # a zero difference: the 1e-8 guard of the backward
    go = torch.randn(shape[0], 1, *shape[2:], generator=gen(92)).to(DEV)
    res = []
    for fused in (False, True):
        a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        out = ofr.channelnorm_diff(a, b) if fused else ofr.ChannelNorm()(a - b)
        res.append((out.detach(),) + torch.autograd.grad(out, (a, b), go))
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)
    assert torch.isfinite(res[1][1]).all()


def test_stn_warp_multi_equals_the_per_scale_loop(ofr):
    """stn_warp_multi(flows, frame) == [stn_warp(f, frame) for f in flows] (models.py:277-281), values and gradients,
    over a pyramid with ragged and tiny scales."""
    B = 3
    frame = torch.rand(B, 1, 40, 36, generator=gen(60)).to(DEV)
    sizes = [(40, 36), (20, 18), (10, 9), (5, 5), (2, 3)]
    flows = [(torch.randn(B, 2, h, w, generator=gen(61 + i)) * 2).to(DEV) for i, (h, w) in enumerate(sizes)]
    gos = [torch.randn(B, 1, h, w, generator=gen(70 + i)).to(DEV) for i, (h, w) in enumerate(sizes)]
    res = []
    for multi in (False, True):
        fl = [f.clone().requires_grad_(True) for f in flows]
        outs = ofr.stn_warp_multi(fl, frame) if multi else [ofr.stn_warp(f, frame) for f in fl]
        total = sum((o * g).sum() for o, g in zip(outs, gos))
        res.append(([o.detach() for o in outs], torch.autograd.grad(total, fl)))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    ref = oracle.stn_warp(flows[1].cpu().double(), frame.cpu().double())
    close(res[1][0][1], ref, atol=1e-5, what="stn multi vs oracle")


def test_registration_loss_equals_stn_then_ofeloss(ofr):
    """registration_loss(flows, moving, fixed) == OFEloss(flows, stn_warp_multi(flows, moving), fixed)
    (models.py:277 + train.py:56): warped images, the four terms and the flow gradients bit for bit, from each of the
    four returned terms, over a pyramid with ragged (VEC=1) and tiny scales; and against the float64 oracle."""
    B = 3
    fixed = torch.rand(B, 1, 40, 36, generator=gen(80)).to(DEV)
    moving = torch.rand(B, 1, 40, 36, generator=gen(81)).to(DEV)
    sizes = [(40, 36), (20, 18), (12, 8), (5, 5), (2, 3)]
    flows = [(torch.randn(B, 2, h, w, generator=gen(82 + i)) * 2).to(DEV) for i, (h, w) in enumerate(sizes)]
    for k in range(4):
        res = []
        for fused in (False, True):
            fl = [f.clone().requires_grad_(True) for f in flows]
            if fused:
                terms, warped = ofr.registration_loss(fl, moving, fixed, lamb_da=0.7, gamma=90.0, zeta=110.0)
                assert not any(w.requires_grad for w in warped)
            else:
                warped = ofr.stn_warp_multi(fl, moving)
                terms = ofr.OFEloss(fl, warped, fixed, lamb_da=0.7, gamma=90.0, zeta=110.0)
            grads = torch.autograd.grad(terms[k], fl, allow_unused=True)
            res.append(([t.detach() for t in terms], [w.detach() for w in warped], grads))
        for a, b in zip(res[0][0], res[1][0]):
            assert torch.equal(a, b), f"term differs (backward from term {k})"
        for a, b in zip(res[0][1], res[1][1]):
            assert torch.equal(a, b), "warped image differs"
        for i, (a, b) in enumerate(zip(res[0][2], res[1][2])):
            if a is None:  # the smoothness term alone does not reach the warp; fused returns zeros + smoothness
                continue
            assert torch.equal(a, b), f"flow gradient {i} differs (backward from term {k})"
    # oracle: stn + OFEloss on the CPU; values against float64, gradients against the reference's own float32 chain
    # (the float32 sampling positions move the warped image by ~3e-6, which the |x|^-0.5 photometric derivative
    # amplifies beyond 1e-4 wherever |fixed - warped| is a few 1e-3: float32 and float64 runs of the REFERENCE differ
    # by 6e-4 on these inputs)
    fd = [f.cpu().double().requires_grad_(True) for f in flows]
    wd = [oracle.stn_warp(f, moving.cpu().double()) for f in fd]
    rres = oracle.ofe_loss(fd, wd, fixed.cpu().double())
    f32 = [f.cpu().requires_grad_(True) for f in flows]
    w32 = [oracle.stn_warp(f, moving.cpu()) for f in f32]
    rg = torch.autograd.grad(oracle.ofe_loss(f32, w32, fixed.cpu())[3], f32)
    fl = [f.clone().requires_grad_(True) for f in flows]
    terms, warped = ofr.registration_loss(fl, moving, fixed)
    gg = torch.autograd.grad(terms[3], fl)
    np.testing.assert_allclose([t.item() for t in terms], [t.item() for t in rres], rtol=2e-6)
    for i, (w, r) in enumerate(zip(warped, wd)):
        close(w, r.detach(), atol=1e-5, what=f"regloss warped[{i}]")
    for i, (g, r) in enumerate(zip(gg, rg)):
        # d/dx (x^2 + 1e-18)^0.25 behaves like |x|^-0.5 near 0: where fixed and warped agree to ~1e-3 the photometric
        # gradient (and with it the flow gradient of that pixel) is ill-conditioned in float32, in the reference too
        h, w = g.shape[2:]
        diff = torch.nn.functional.interpolate(fixed, (h, w), mode="bilinear", align_corners=False) - warped[i]
        keep = (diff.abs() > 1e-3).double().cpu()
        close(g.double().cpu() * keep, r.double() * keep, atol=1e-6, rtol=1e-4, what=f"regloss gflow[{i}]")


def test_registration_loss_vs_reference_golden(ofr, golden):
    """The fused kernels against the reference's own chain stn -> OFEloss run on the same inputs (regloss.npz, generated
    by tests/golden/make_golden.py from models.py:256-268,277 and loss.py:66-84): warped images, the four terms, and
    d total / d flow away from the pixels where the photometric derivative is ill-conditioned in float32."""
    g = golden("regloss")
    fixed, moving = torch.from_numpy(g["fixed"]).to(DEV), torch.from_numpy(g["moving"]).to(DEV)
    flows = [torch.from_numpy(g[f"flow{i}"]).to(DEV).requires_grad_(True) for i in range(4)]
    terms, warped = ofr.registration_loss(flows, moving, fixed)
    np.testing.assert_allclose([t.item() for t in terms], g["ofe"], rtol=2e-6)
    grads = torch.autograd.grad(terms[3], flows)
    for i in range(4):
        close(warped[i], torch.from_numpy(g[f"warped{i}"]), atol=1e-6, what=f"regloss golden warped{i}")
        h, w = flows[i].shape[2:]
        diff = torch.nn.functional.interpolate(fixed, (h, w), mode="bilinear", align_corners=False) - warped[i]
        keep = (diff.abs() > 1e-3).double().cpu()
        close(grads[i].double().cpu() * keep, torch.from_numpy(g[f"gflow{i}"]).double() * keep, atol=1e-6, rtol=1e-4,
              what=f"regloss golden gflow{i}")


def test_registration_loss_argument_errors(ofr):
    fixed = torch.rand(2, 1, 16, 16).to(DEV)
    flow = torch.zeros(2, 2, 16, 16).to(DEV)
    with pytest.raises(ValueError):
        ofr.registration_loss([flow], fixed.clone().requires_grad_(True), fixed)
    with pytest.raises(ValueError):
        ofr.registration_loss([], fixed, fixed)
    with pytest.raises(ValueError):
        ofr.registration_loss([flow[:, :1]], fixed, fixed)
    # zero flow: the warped image is the stn identity resampling of the moving image, x * (W - 1) / W
    terms, warped = ofr.registration_loss([flow], fixed, fixed)
    close(warped[0], oracle.stn_warp(flow.cpu(), fixed.cpu()), atol=1e-6, what="regloss zero flow")


# =============================================================================== loss
def test_loss_vs_reference_golden(ofr, golden):
    g = golden("loss")
    fixed = torch.from_numpy(g["fixed"]).to(DEV)
    flows = [torch.from_numpy(g[f"flow{i}"]).to(DEV).requires_grad_(True) for i in range(4)]
    warped = [torch.from_numpy(g[f"warped{i}"]).to(DEV).requires_grad_(True) for i in range(4)]
    res = ofr.OFEloss(flows, warped, fixed)
    assert all(t.dtype == torch.float64 and t.dim() == 0 for t in res)
    np.testing.assert_allclose([t.item() for t in res], g["ofe"], rtol=2e-6)
    grads = torch.autograd.grad(res[3], flows + warped)
    for i in range(4):
        close(grads[i], torch.from_numpy(g[f"gflow{i}"]), atol=1e-6, rtol=1e-4, what=f"loss gflow{i}")
        close(grads[4 + i], torch.from_numpy(g[f"gwarped{i}"]), atol=1e-6, rtol=1e-4, what=f"loss gwarped{i}")
    res = ofr.OFEloss(flows[:2], warped[:2], fixed, lamb_da=0.25, gamma=10.0, zeta=3.0)
    np.testing.assert_allclose([t.item() for t in res], g["ofe_alt"], rtol=2e-6)
    for i in range(4):
        np.testing.assert_allclose(ofr.photometric_loss(fixed, warped[i].detach()).item(), g[f"photo{i}"], rtol=2e-6)
        np.testing.assert_allclose(ofr.correlation_loss(fixed, warped[i].detach()).item(), g[f"ncc{i}"], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(ofr.smoothness_loss(flows[i].detach()).item(), g[f"smooth{i}"], rtol=2e-6)


def test_loss_individual_term_gradients(ofr):
    fixed = torch.rand(2, 1, 64, 64, generator=gen(20))
    sizes = [64, 32, 16, 8, 4]
    flows = [torch.randn(2, 2, s, s, generator=gen(21 + s)) for s in sizes]
    warps = [torch.rand(2, 1, s, s, generator=gen(31 + s)) for s in sizes]
    for k in range(4):  # backward from each of the four returned terms
        cf = [f.clone().requires_grad_(True) for f in flows]
        cw = [w.clone().requires_grad_(True) for w in warps]
        ref = oracle.ofe_loss(cf, cw, fixed)
        rg = torch.autograd.grad(ref[k], cf + cw, allow_unused=True)
        df = [f.to(DEV).requires_grad_(True) for f in flows]
        dw = [w.to(DEV).requires_grad_(True) for w in warps]
        got = ofr.OFEloss(df, dw, fixed.to(DEV))
        np.testing.assert_allclose(got[k].item(), ref[k].item(), rtol=2e-6)
        dg = torch.autograd.grad(got[k], df + dw, allow_unused=True)
        for i, (a, b) in enumerate(zip(dg, rg)):
            if b is None:
                assert a is None or float(a.abs().max()) == 0.0
            else:
                close(a, b, atol=1e-6, rtol=1e-4, what=f"term{k} grad{i}")


def test_loss_degenerate_and_reproducible(ofr, golden):
    g = golden("loss")
    fixed = torch.from_numpy(g["fixed"]).to(DEV)
    const = torch.full((3, 1, 16, 16), 0.5, device=DEV)
    assert ofr.correlation_loss(fixed, const).item() == float(g["ncc_const"]) == 0.0   # loss.py:59-60
    z = ofr.smoothness_loss(torch.zeros(3, 2, 8, 8, device=DEV)).item()
    np.testing.assert_allclose(z, float(g["smooth_zero"]), rtol=1e-6)
    flows = [torch.randn(3, 2, 32, 32, generator=gen(40)).to(DEV)]
    warps = [torch.rand(3, 1, 32, 32, generator=gen(41)).to(DEV)]
    a = ofr.OFEloss(flows, warps, fixed)
    b = ofr.OFEloss(flows, warps, fixed)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_ncc_constant_image_rule(ofr, golden):
    """NCC of constant warped images: the reference (ncc_const.npz) returns 0 when the float32 mean of the constant is
    exact and ~1 otherwise (0.7); the kernels return 0 -- corr = 1, zero gradient -- for EVERY constant image
    (documented divergence on a pathological input, include/ofr_b200.h)."""
    g = golden("ncc_const")
    fixed = torch.from_numpy(g["fixed"]).to(DEV)
    for v, want in zip(g["values"], g["ref_loss"]):
        w = torch.full((3, 1, 16, 16), float(v), device=DEV, requires_grad=True)
        got = ofr.correlation_loss(fixed, w)
        assert got.dtype == torch.float32 and got.dim() == 0     # as loss.py's leaf functions return
        assert got.item() == 0.0
        if want == 0.0:
            assert got.item() == want
        (gw,) = torch.autograd.grad(got, w)
        assert float(gw.abs().max()) == 0.0


def test_loss_accepts_misaligned_views_and_huge_batches(ofr):
    """(i) contiguous views whose storage offset is not a multiple of 16 bytes must take the scalar path, not
    fault; (ii) more (image, strip) items than the per-scale CTA cap are walked by a loop, not refused."""
    B, S = 2, 32
    gcpu = gen(120)
    fixed = torch.rand(B, 1, S, S, generator=gcpu).to(DEV)
    moving = torch.rand(B, 1, S, S, generator=gcpu).to(DEV)
    sizes = [(32, 32), (16, 16), (8, 8)]
    flows = [(torch.randn(B, 2, h, w, generator=gcpu) * 1.5).to(DEV) for h, w in sizes]
    warps = [torch.rand(B, 1, h, w, generator=gcpu).to(DEV) for h, w in sizes]

    def shifted(t):  # same values in a buffer whose data pointer is 4 bytes past a 16-byte boundary
        buf = torch.empty(t.numel() + 1, device=DEV)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.is_contiguous() and v.data_ptr() % 16 == 4
        return v

    def run(fl, wp, fx, mv):
        fl = [f.detach().requires_grad_(True) for f in fl]
        wp = [w.detach().requires_grad_(True) for w in wp]
        res = ofr.OFEloss(fl, wp, fx)
        g1 = torch.autograd.grad(res[3], fl + wp)
        fl2 = [f.detach().requires_grad_(True) for f in fl]
        terms, warped = ofr.registration_loss(fl2, mv, fx)
        g2 = torch.autograd.grad(terms[3], fl2)
        return list(res) + list(g1) + list(terms) + list(warped) + list(g2)

    a = run(flows, warps, fixed, moving)
    b = run([shifted(f) for f in flows], [shifted(w) for w in warps], shifted(fixed), shifted(moving))
    for i, (x, y) in enumerate(zip(a, b)):
        close(y, x, atol=1e-6, rtol=1e-6, what=f"misaligned loss item {i}")
    # (ii) 9000 images x 1 strip > 8192 CTAs per scale
    Bb = 9000
    fx = torch.rand(Bb, 1, 8, 8, generator=gcpu).to(DEV)
    fl = [torch.randn(Bb, 2, 8, 8, generator=gcpu).to(DEV).requires_grad_(True)]
    wp = [torch.rand(Bb, 1, 8, 8, generator=gcpu).to(DEV).requires_grad_(True)]
    res = ofr.OFEloss(fl, wp, fx)
    gg = torch.autograd.grad(res[3], fl + wp)
    ref = oracle.ofe_loss([f.detach().cpu().double().requires_grad_(True) for f in fl],
                          [w.detach().cpu().double().requires_grad_(True) for w in wp], fx.cpu().double())
    np.testing.assert_allclose([float(t) for t in res], [float(t) for t in ref], rtol=2e-6)
    assert all(bool(torch.isfinite(g_).all()) for g_ in gg)


# =============================================================================== model level
def test_pwcnet_training_step_matches_oracle_layers(ofr):
    """The whole hot path in situ: host PWC-Net + stn warps + OFEloss, CUDA layers on the GPU against the same model
    with every hot-path layer served by the CPU oracle (same weights, same input)."""
    from ofr_b200.harness import PWCNetHost, RegistrationStep, synthetic_pairs

    class OracleCorrelation(torch.nn.Module):
        def forward(self, a, b):
            return oracle.correlation(a, b, 4, 1, 4, 1, 1)

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # float32 convolutions on both sides
    try:
        torch.manual_seed(6)
        cpu_host = PWCNetHost(md=4, corr=OracleCorrelation(), warp=oracle.pwc_warp)
        gpu_host = PWCNetHost(md=4).to(DEV)
        gpu_host.load_state_dict(cpu_host.state_dict())
        cpu_step = RegistrationStep(cpu_host, oracle.stn_warp, oracle.ofe_loss)
        gpu_step = RegistrationStep(gpu_host, ofr.stn_warp, ofr.OFEloss)
        pair = synthetic_pairs(2, 128, seed=6)
        ref = cpu_step(pair)
        got = gpu_step(pair.to(DEV))
        for r, g_ in zip(ref, got):
            assert abs(g_.item() - r.item()) <= 2e-3 * abs(r.item()) + 1e-6, (g_.item(), r.item())
        ref[3].backward()
        got[3].backward()
        for name in ["conv6b.0.weight", "conv2_0.0.weight", "predict_flow2.weight", "conv1a.0.weight"]:
            a = dict(gpu_host.named_parameters())[name].grad.cpu().double()
            b = dict(cpu_host.named_parameters())[name].grad.double()
            rel = (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)
            assert rel < 2e-2, (name, rel)
    finally:
        torch.backends.cudnn.allow_tf32 = old


# =============================================================================== no out-of-bounds writes
class _Guarded:
    """A float32 device buffer of `n` elements framed by two guard zones filled with a sentinel; every output and
    workspace of a raw C-ABI call is allocated through it, and `check()` asserts the guards are untouched."""
    PAD = 4096  # elements on each side (16 KB), keeps the payload 256-byte aligned

    def __init__(self, n, dtype=torch.float32):
        self.n, self.dtype = int(n), dtype
        self.buf = torch.full((self.n + 2 * self.PAD,), 12345.0, dtype=torch.float32, device=DEV).to(dtype)
        self.sentinel = self.buf[0].clone()

    @property
    def t(self):
        return self.buf[self.PAD:self.PAD + self.n]

    def ptr(self):
        import ctypes
        return ctypes.c_void_p(self.t.data_ptr())

    def check(self, what):
        lo, hi = self.buf[:self.PAD], self.buf[self.PAD + self.n:]
        assert bool((lo == self.sentinel).all()) and bool((hi == self.sentinel).all()), f"{what}: guard zone overwritten"


@pytest.mark.parametrize("shape", [(2, 13, 12, 20), (1, 8, 6, 136), (2, 196, 4, 4), (3, 40, 16, 16), (1, 32, 30, 36), (2, 33, 50, 72)])
def test_correlation_writes_stay_inside_their_buffers(ofr, shape):
    """Raw C-ABI calls with guard zones around every written buffer (TMA kernels, cluster split, ragged tiles)."""
    import ctypes
    from ofr_b200 import _lib
    L = _lib.lib()
    B, C, H, W = shape
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f1 = torch.randn(B, C, H, W, generator=gen(40)).to(DEV)
    f2 = torch.randn(B, C, H, W, generator=gen(41)).to(DEV)
    out, g1, g2 = _Guarded(B * 81 * H * W), _Guarded(B * C * H * W), _Guarded(B * C * H * W)
    assert L.ofr_corr_fwd_act(P(f1), P(f2), out.ptr(), B, C, H, W, 4, 1, 4, 1, 1, 0.1, None, 0, st) == 0
    go = torch.randn(B, 81, H, W, generator=gen(42)).to(DEV)
    assert L.ofr_corr_bwd(P(f1), P(f2), P(go), g1.ptr(), g2.ptr(), B, C, H, W, 4, 1, 4, 1, 1, None, 0, st) == 0
    torch.cuda.synchronize()
    for name, gbuf in (("corr out", out), ("corr gI1", g1), ("corr gI2", g2)):
        gbuf.check(name)
        assert bool(torch.isfinite(gbuf.t).all()) and not bool((gbuf.t == 12345.0).any()), f"{name}: unwritten elements"
    ref = torch.nn.functional.leaky_relu(oracle.correlation(f1.cpu().double(), f2.cpu().double(), 4, 1, 4, 1, 1), 0.1)
    close(out.t.view(B, 81, H, W), ref, atol=1e-5, what="guarded corr fwd")


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(2, 5, 9, 13), (1, 24, 32, 32), (2, 1, 40, 36), (1, 130, 8, 8)])
def test_warp_writes_stay_inside_their_buffers(ofr, mode, shape):
    import ctypes
    from ofr_b200 import _lib
    L = _lib.lib()
    B, C, H, W = shape
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = torch.randn(B, C, H, W, generator=gen(43)).to(DEV)
    fl = (torch.randn(B, 2, H, W, generator=gen(44)) * 3).to(DEV)
    go = torch.randn(B, C, H, W, generator=gen(45)).to(DEV)
    out, g1, g2 = _Guarded(B * C * H * W), _Guarded(B * C * H * W), _Guarded(B * 2 * H * W)
    nws = int(L.ofr_resample_bwd_workspace(B, C, H, W, H, W, mode))
    ws = _Guarded((nws + 3) // 4)
    assert L.ofr_resample_fwd_scaled(P(x), P(fl), out.ptr(), B, C, H, W, H, W, mode, 1.25, st) == 0
    assert L.ofr_resample_bwd_scaled(P(x), P(fl), P(go), g1.ptr(), g2.ptr(), B, C, H, W, H, W, mode, 1.25, ws.ptr(),
                                     nws, st) == 0
    g2b = _Guarded(B * 2 * H * W)  # flow gradient only: the other kernel family
    assert L.ofr_resample_bwd_scaled(P(x), P(fl), P(go), None, g2b.ptr(), B, C, H, W, H, W, mode, 1.25, None, 0, st) == 0
    torch.cuda.synchronize()
    for name, gbuf in (("warp out", out), ("warp gI1", g1), ("warp gI2", g2), ("warp workspace", ws), ("warp gI2'", g2b)):
        gbuf.check(f"{name} mode {mode}")
    for name, gbuf in (("warp out", out), ("warp gI1", g1), ("warp gI2", g2), ("warp gI2'", g2b)):
        assert not bool((gbuf.t == 12345.0).any()), f"{name}: unwritten elements"
    close(g2b.t, g2.t, atol=1e-5, rtol=1e-4, what="flow gradient: fused kernel vs flow-only kernel")


def test_loss_and_channelnorm_writes_stay_inside_their_buffers(ofr):
    import ctypes
    from ofr_b200 import _lib
    L = _lib.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    B, S = 3, 40
    sizes = [(40, 40), (20, 20), (10, 12), (5, 7)]
    fixed = torch.rand(B, 1, S, S, generator=gen(46)).to(DEV)
    flows = [torch.randn(B, 2, h, w, generator=gen(47)).to(DEV) for h, w in sizes]
    warps = [torch.rand(B, 1, h, w, generator=gen(48)).to(DEV) for h, w in sizes]
    gfl = [_Guarded(B * 2 * h * w) for h, w in sizes]
    gwp = [_Guarded(B * h * w) for h, w in sizes]
    n = len(sizes)
    arr = (_lib.LossScale * n)()
    for i, (h, w) in enumerate(sizes):
        arr[i].warped, arr[i].flow = warps[i].data_ptr(), flows[i].data_ptr()
        arr[i].g_warped, arr[i].g_flow = gwp[i].t.data_ptr(), gfl[i].t.data_ptr()
        arr[i].h, arr[i].w = h, w
    out4, state = _Guarded(4, torch.float64), _Guarded(n * 16, torch.float64)
    nws = int(L.ofr_loss_workspace(n))
    ws = _Guarded((nws + 3) // 4)
    g4 = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64, device=DEV)
    assert L.ofr_loss_fwd(P(fixed), B, S, S, arr, n, 0.5, 100.0, 100.0, out4.ptr(), state.ptr(), ws.ptr(), nws, st) == 0
    assert L.ofr_loss_bwd(P(fixed), B, S, S, arr, n, 0.5, 100.0, 100.0, P(g4), state.ptr(), st) == 0
    x = torch.randn(2, 3, 7, 9, generator=gen(49)).to(DEV)
    o, gi = _Guarded(2 * 7 * 9), _Guarded(x.numel())
    gout = torch.randn(2, 1, 7, 9, generator=gen(50)).to(DEV)
    assert L.ofr_chnorm_fwd(P(x), o.ptr(), 2, 3, 7, 9, st) == 0
    assert L.ofr_chnorm_bwd(P(x), o.ptr(), P(gout), gi.ptr(), 2, 3, 7, 9, st) == 0
    torch.cuda.synchronize()
    for i in range(n):
        gfl[i].check(f"loss g_flow[{i}]")
        gwp[i].check(f"loss g_warped[{i}]")
        assert not bool((gfl[i].t == 12345.0).any()) and not bool((gwp[i].t == 12345.0).any())
    for name, gbuf in (("loss out4", out4), ("loss state", state), ("loss workspace", ws), ("chnorm out", o), ("chnorm gI", gi)):
        gbuf.check(name)
    ref = oracle.ofe_loss([f.cpu() for f in flows], [w.cpu() for w in warps], fixed.cpu())
    np.testing.assert_allclose(out4.t.cpu().numpy(), [float(r) for r in ref], rtol=2e-6)


# =============================================================================== BASELINE sizes (B = 24, S = 512)
# The CPU oracle takes minutes at these sizes, so the kernels are checked through size-independent properties and
# against the oracle's own (device-agnostic) PyTorch code run on the GPU in float64.
FULL_B = 24


@pytest.mark.parametrize("C,h", [(32, 128), (64, 64), (96, 32), (128, 16), (196, 8)])
def test_fullsize_correlation_properties(ofr, C, h):
    """PWC-Net pyramid shapes at batch 24: (i) one-hot features pick the shifted pixel, (ii) the operator is bilinear,
    so <dL/df1, f1> == <gO, out> == <dL/df2, f2> (Euler), (iii) forward vs the oracle code on the GPU in float64."""
    B = FULL_B
    g0 = torch.Generator(device=DEV).manual_seed(80)
    f1 = torch.randn(B, C, h, h, device=DEV, generator=g0, requires_grad=True)
    f2 = torch.randn(B, C, h, h, device=DEV, generator=g0, requires_grad=True)
    corr = ofr.Correlation(4, 1, 4, 1, 1, 1)
    out = corr(f1, f2)
    go = torch.randn(out.shape, device=DEV, generator=g0)
    g1, g2 = torch.autograd.grad(out, (f1, f2), go)
    lhs = (go.double() * out.detach().double()).sum().item()
    for name, g, f in (("f1", g1, f1), ("f2", g2, f2)):
        rhs = (g.double() * f.detach().double()).sum().item()
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (name, lhs, rhs)
    ref = oracle.correlation(f1.detach().double(), f2.detach().double(), 4, 1, 4, 1, 1)
    close(out, ref, atol=1e-5, what=f"full-size corr C{C} {h}x{h}")
    # one-hot: f1 = e_c at every pixel, f2 = e_c * ramp  ->  out[b, (tj+4)*9+(ti+4), y, x] = ramp[y+tj, x+ti] / C
    ramp = torch.arange(h * h, device=DEV, dtype=torch.float32).view(1, 1, h, h) / (h * h)
    a = torch.zeros(1, C, h, h, device=DEV)
    bb = torch.zeros(1, C, h, h, device=DEV)
    a[:, C // 2] = 1.0
    bb[:, C // 2] = ramp[0, 0]
    o = corr(a, bb)
    ys, xs = (1, 2) if h > 4 else (0, 0)
    for tj, ti in ((-4, -4), (0, 0), (3, -2), (4, 4)):
        y, x = ys + tj, xs + ti
        want = ramp[0, 0, y, x].item() / C if 0 <= y < h and 0 <= x < h else 0.0
        assert abs(o[0, (tj + 4) * 9 + (ti + 4), ys, xs].item() - want) <= 1e-7


@pytest.mark.parametrize("C,h", [(32, 128), (64, 64), (96, 32), (128, 16)])
def test_fullsize_pwc_warp_properties(ofr, C, h):
    """PWC warp at batch 24: linear in the image, so <dL/dx, x> == <gO, out>; forward and both gradients vs the oracle
    code on the GPU in float64."""
    B = FULL_B
    g0 = torch.Generator(device=DEV).manual_seed(81)
    x = torch.randn(B, C, h, h, device=DEV, generator=g0, requires_grad=True)
    coarse = torch.randn(B, 2, max(h // 8, 1), max(h // 8, 1), device=DEV, generator=g0) * 2
    fl = torch.nn.functional.interpolate(coarse, size=(h, h), mode="bilinear", align_corners=False).requires_grad_(True)
    out = ofr.pwc_warp(x, fl)
    go = torch.randn(out.shape, device=DEV, generator=g0)
    gx, gf = torch.autograd.grad(out, (x, fl), go)
    lhs = (go.double() * out.detach().double()).sum().item()
    rhs = (gx.double() * x.detach().double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)
    # the reference computes the sampling positions in float32: at this size a float64 oracle differs from IT by more
    # than the tolerance, so the comparison is with the oracle in the reference's own arithmetic (float32, CPU)
    xd, fd = x.detach().cpu().requires_grad_(True), fl.detach().cpu().requires_grad_(True)
    ref = oracle.pwc_warp(xd, fd)
    rgx, rgf = torch.autograd.grad(ref, (xd, fd), go.cpu())
    close(out, ref, atol=1e-5, what="full-size pwc warp")
    close(gx, rgx, atol=1e-6, rtol=1e-4, what="full-size pwc warp dImg")
    close(gf, rgf, atol=1e-5, rtol=1e-4, what="full-size pwc warp dFlow")


def test_fullsize_stn_channelnorm_loss(ofr):
    """512x512, batch 24: stn pyramid + fused 7-scale loss vs the oracle code on the GPU (float64), ChannelNorm vs
    torch, and a run-to-run bitwise check of the whole chain."""
    B, S = FULL_B, 512
    g0 = torch.Generator(device=DEV).manual_seed(82)
    fixed = torch.rand(B, 1, S, S, device=DEV, generator=g0)
    moving = torch.rand(B, 1, S, S, device=DEV, generator=g0)
    flows = [(torch.randn(B, 2, S >> i, S >> i, device=DEV, generator=g0) * 2).requires_grad_(True) for i in range(7)]

    def run():
        warped = ofr.stn_warp_multi(flows, moving)
        res = ofr.OFEloss(flows, warped, fixed)
        return res, torch.autograd.grad(res[3], flows), warped

    res, grads, warped = run()
    res2, grads2, _ = run()
    assert all(torch.equal(a, b) for a, b in zip(res, res2)) and all(torch.equal(a, b) for a, b in zip(grads, grads2))
    # stn in the reference's own arithmetic (float32, CPU: at 512 px a float64 oracle is 1e-5 away from the reference
    # itself); the loss, given the same warped images, against the oracle code in float64 on the GPU
    for f, w in zip(flows[:3], warped[:3]):
        close(w, oracle.stn_warp(f.detach().cpu(), moving.cpu()), atol=1e-5, what="full-size stn")
    fd = [f.detach().double().requires_grad_(True) for f in flows]
    wd = [w.detach().double().requires_grad_(True) for w in warped]
    rres = oracle.ofe_loss(fd, wd, fixed.double())
    rg = torch.autograd.grad(rres[3], fd + wd)
    wl = [w.detach().requires_grad_(True) for w in warped]
    lres = ofr.OFEloss(flows, wl, fixed)
    lg = torch.autograd.grad(lres[3], flows + wl)
    np.testing.assert_allclose([float(t.detach()) for t in lres], [float(t.detach()) for t in rres], rtol=2e-6)
    np.testing.assert_allclose([float(t.detach()) for t in res], [float(t.detach()) for t in rres], rtol=2e-6)
    n = len(flows)
    for i, (g, r) in enumerate(zip(lg, rg)):
        if i >= n:
            # d/dx (x^2 + 1e-18)^0.25 behaves like |x|^-0.5 near 0: where fixed and warped agree to ~1e-3 the photometric
            # gradient is ill-conditioned in float32 (in the reference too) -> compare away from those pixels
            h, w = g.shape[2:]
            diff = torch.nn.functional.interpolate(fixed, (h, w), mode="bilinear", align_corners=False) - warped[i - n]
            keep = (diff.abs() > 1e-3).double()
            g, r = g.double() * keep, r * keep
        close(g, r, atol=1e-7, rtol=1e-4, what=f"full-size loss grad[{i}]")
    x = torch.randn(B, 2, S, S, device=DEV, generator=g0, requires_grad=True)
    out = ofr.ChannelNorm()(x)
    gout = torch.randn(out.shape, device=DEV, generator=g0)
    (gx,) = torch.autograd.grad(out, x, gout)
    ref = x.detach().double().pow(2).sum(1, keepdim=True).sqrt()
    close(out, ref, atol=1e-6, what="full-size channelnorm")
    close(gx, gout.double() * x.detach().double() / (ref + 1e-8), atol=1e-6, rtol=1e-5, what="full-size channelnorm grad")

"""CPU: the measurement harness (host model, synthetic data, training step, data-parallel sharding).

The harness takes its hot-path layers by injection; here they are the CPU oracle, so the host-side logic
(model wiring, step, DDP gradient averaging over gloo with world_size 2) is covered without a GPU.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle

REF = "/root/reference"


class OracleCorrelation(torch.nn.Module):
    def forward(self, a, b):
        return oracle.correlation(a, b, 4, 1, 4, 1, 1)


def make_step(seed=6):
    from ofr_b200.harness import PWCNetHost, RegistrationStep
    torch.manual_seed(seed)
    host = PWCNetHost(md=4, corr=OracleCorrelation(), warp=oracle.pwc_warp)
    host.deconv0.requires_grad_(False)
    return RegistrationStep(host, oracle.stn_warp, oracle.ofe_loss)


def test_host_model_shapes_and_parameter_count():
    step = make_step()
    assert sum(p.numel() for p in step.predictor.parameters()) == 9374184
    assert len(step.predictor.state_dict()) == len(list(step.predictor.parameters()))  # no buffers
    x = torch.rand(1, 2, 64, 64)
    with torch.no_grad():
        flows = step.predictor(x)
    assert [tuple(f.shape) for f in flows] == [(1, 2, s, s) for s in (64, 32, 16, 8, 4, 2, 1)]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_host_model_equals_reference_class():
    """Same state_dict keys/shapes as PWC/models/PWCNet.py:38-135 and bitwise-equal outputs with shared weights."""
    import ofr_b200
    ofr_b200.install_shims()
    sys.path.insert(0, os.path.join(REF, "PWC", "models"))
    real_cuda = torch.Tensor.cuda
    try:
        import PWCNet
        ref = PWCNet.PWCDCNet(md=4)
        ref.corr = OracleCorrelation()
        host = make_step().predictor
        sd = ref.state_dict()
        assert set(sd) == set(host.state_dict())
        host.load_state_dict(sd)
        torch.Tensor.cuda = lambda self, *a, **k: self  # PWCNet.py:169 moves the mask unconditionally
        x = torch.rand(1, 2, 64, 64, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            a, b = ref(x), host(x)
        for u, v in zip(a, b):
            assert torch.equal(u, v)
    finally:
        torch.Tensor.cuda = real_cuda
        sys.path.remove(os.path.join(REF, "PWC", "models"))
        sys.modules.pop("PWCNet", None)
        ofr_b200.uninstall_shims()


def test_synthetic_pairs_are_deterministic_and_normalised():
    from ofr_b200.harness import synthetic_pairs
    a, b = synthetic_pairs(2, 64, seed=6), synthetic_pairs(2, 64, seed=6)
    assert torch.equal(a, b) and a.shape == (2, 2, 64, 64)
    assert float(a[:, 0].amin()) == 0.0 and abs(float(a[:, 0].amax()) - 1.0) < 1e-6
    assert not torch.equal(a, synthetic_pairs(2, 64, seed=7))
    assert float((a[:, 0] - a[:, 1]).abs().mean()) > 1e-3  # moving really is deformed


def test_training_step_reduces_loss():
    from ofr_b200.harness import synthetic_pairs, make_optimizer, train_step
    step = make_step()
    opt = make_optimizer(step.predictor, lr=1e-4)
    pair = synthetic_pairs(1, 64, seed=6)
    losses = [float(train_step(step, opt, pair)) for _ in range(3)]
    assert all(l == l and l > 0 for l in losses)
    assert losses[-1] < losses[0]


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ofr_b200.harness import synthetic_pairs
        torch.set_num_threads(2)
        step = make_step(seed=6)  # identical weights on every rank
        local = make_step(seed=6)
        step.predictor = torch.nn.parallel.DistributedDataParallel(step.predictor)
        pair = synthetic_pairs(1, 64, seed=10 + rank)  # rank-offset data shard
        step(pair)[3].backward()
        local(pair)[3].backward()
        mine = torch.cat([p.grad.flatten() for p in local.predictor.parameters() if p.grad is not None]).double()
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = sum(gathered) / world  # DDP averages the per-shard gradients
        got = torch.cat([p.grad.flatten() for p in step.predictor.parameters() if p.grad is not None]).double()
        q.put((rank, float((got - expect).abs().max()), float(expect.abs().max())))
    finally:
        dist.destroy_process_group()


def test_ddp_gloo_world2_averages_shard_gradients():
    """Batch sharding needs no collective inside the layers: one all-reduce of the gradients per step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(2)]
    for rank, err, scale in res:
        assert err <= 1e-5 * max(scale, 1.0), (rank, err, scale)


# ----------------------------------------------------------------------------------------- round 2: reference in situ
def test_staged_reference_is_byte_identical():
    """tools/stage_reference.py: baseline/_ref holds byte-for-byte copies (SHA-256 manifest) of the reference's model
    files; where /root/reference exists the staged files are compared with it too."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    try:
        import stage_reference
        if stage_reference.stage() is None:
            pytest.skip("no reference tree and nothing staged on this box")
        assert stage_reference.check() == []
    finally:
        sys.path.pop(0)


def test_reference_flownet2_runs_unmodified_with_injected_layers():
    """harness/reference_env.py: the reference's own ``flownet2/models.py`` FlowNet2 and ``loss.py`` / ``models.py`` import
    and run with stand-in layers at the import paths (here the oracle's, on the CPU; on the GPU box the CUDA layers:
    tests/test_gpu_reference_models.py), and every trace is removed from sys.modules afterwards."""
    from ofr_b200.harness import reference_env
    if reference_env.reference_dir() is None:
        pytest.skip("reference files not staged")
    before = set(sys.modules)
    with reference_env.ReferenceEnv(oracle.MODULES) as e:
        torch.manual_seed(6)
        net = e.flownet2()
        assert sum(p.numel() for p in net.parameters()) == 162493554
        kinds = {type(m) for m in net.modules()}
        assert set(oracle.MODULES.values()) <= kinds
        with torch.no_grad():
            out = net(torch.rand(1, 2, 128, 128))     # (64 px would leave BatchNorm one value per channel)
        assert out[0].shape == (1, 2, 128, 128) and out[0] is out[1]
        stn = e.models_module().opticalFlowReg.stn
        fl = torch.randn(1, 2, 16, 16)
        fr = torch.rand(1, 1, 32, 32)
        assert torch.allclose(stn(None, fl, fr), oracle.stn_warp(fl, fr), atol=1e-6)
        assert callable(e.loss_module().OFEloss)
    leaked = [m for m in set(sys.modules) - before if m.split(".")[0] in
              ("flownet2", "networks", "models", "loss", "utils", "PWCNet", "monai", "matplotlib", "correlation_package")]
    assert leaked == []


def test_registration_wrapper_helpers():
    """registration.py: the cached line grid equals utils.py:15-23, label rounding equals models.py:286 (np.rint + clip)."""
    import numpy as np
    import ofr_b200
    g = ofr_b200.grid_generator(256)
    want = torch.zeros(256, 256)
    for i in range(7, 255, 16):
        want[i, :] = 1.0
        want[:, i] = 1.0
    assert torch.equal(g[0, 0], want) and ofr_b200.grid_generator(256) is g
    x = torch.tensor([[-0.7, 0.5, 1.5, 2.5, 2.51, 3.49, 3.5, 9.0]]).view(1, 1, 1, 8)
    got = ofr_b200.round_clamp_labels(x)
    assert torch.equal(got, torch.from_numpy(np.clip(np.rint(x.numpy()), 0, 3)))


def test_elastic_generator_control_grid_rule():
    """pairs.py: MONAI's create_control_grid size rule (spacing 16 + border) and that the bicubic upsampling of the
    control grid covers the slice after the centre crop."""
    import ofr_b200
    assert ofr_b200.control_grid_size(256, 16) == 20 and ofr_b200.control_grid_size(512, 16) == 36
    for size in (64, 256, 255, 512):
        n = ofr_b200.control_grid_size(size, 16)
        assert n * 16 >= size + 2 * 16


def test_bench_reference_arm_is_same_config_by_default():
    """bench.py --impl reference runs the GPU arm's batch unless --cpu-batch says otherwise."""
    import importlib
    bench = importlib.import_module("bench")
    src = open(bench.__file__).read()
    assert '"--cpu-batch", type=int, default=0' in src and '"same_config": batch == args.batch' in src


def test_flow_heads_reclass_in_place_and_fall_back_on_cpu():
    """``accelerate_flow_heads`` on the CPU: only the two-output-channel 3x3 / 4x4-stride-2 layers change class,
    parameters and state-dict keys stay, and a CPU tensor goes through the parent ``nn`` forward (the kernels are CUDA-only;
    this is the reference's own op, not a fallback of the product path)."""
    import torch.nn as nn
    import ofr_b200
    from ofr_b200.harness import PWCNetHost
    torch.manual_seed(0)
    net = PWCNetHost(md=4, fast_heads=False)
    keys = list(net.state_dict().keys())
    before = {k: v.clone() for k, v in net.state_dict().items()}
    changed = ofr_b200.accelerate_flow_heads(net)
    assert sorted(changed) == sorted([f"predict_flow{i}" for i in (6, 5, 4, 3, 2)] + ["dc_conv7"] +
                                     [f"deconv{i}" for i in (6, 5, 4, 3, 2, 1)] + [f"upfeat{i}" for i in (6, 5, 4, 3)])
    assert type(net.deconv0) is nn.ConvTranspose2d and type(net.conv1a[0]) is nn.Conv2d
    assert isinstance(net.predict_flow2, ofr_b200.FlowConv2d) and isinstance(net.upfeat3, ofr_b200.FlowUpsample2d)
    assert list(net.state_dict().keys()) == keys
    assert all(torch.equal(v, before[k]) for k, v in net.state_dict().items())
    assert ofr_b200.accelerate_flow_heads(net) == []                      # idempotent
    x = torch.randn(2, net.predict_flow6.in_channels, 4, 4)
    assert torch.equal(net.predict_flow6(x), torch.nn.functional.conv2d(x, net.predict_flow6.weight, net.predict_flow6.bias, 1, 1))
    f = torch.randn(2, 2, 4, 4)
    assert torch.equal(net.deconv6(f), torch.nn.functional.conv_transpose2d(f, net.deconv6.weight, net.deconv6.bias, 2, 1))
    fresh = PWCNetHost(md=4)                                              # product default: heads re-classed at construction
    assert isinstance(fresh.dc_conv7, ofr_b200.FlowConv2d)
    fresh.load_state_dict(net.state_dict())

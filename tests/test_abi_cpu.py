"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares;
the host-side layer API validates its arguments and refuses to run without CUDA tensors."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.fixture(scope="module")
def built(ofr):
    ofr.build()
    return ofr


def test_header_symbols_all_exported(built):
    header = open(os.path.join(ROOT, "include", "ofr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ofr_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(built.EXPORTS), (declared, built.EXPORTS)
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libofr_b200.so does not export {name}"


def test_abi_version_and_error_strings(built):
    lib = built._lib.lib()
    assert lib.ofr_abi_version() == 3
    assert lib.ofr_error_string(0) == b"ok"
    for code in (1, 2, 3):
        assert len(lib.ofr_error_string(code)) > 5
    assert lib.ofr_error_string(77) == b"unknown ofr error code"


def test_host_side_argument_checks_need_no_gpu(built):
    lib = built._lib.lib()
    # shape helper is pure host code
    assert built.correlation_out_shape(32, 32, 20, 1, 20, 1, 2) == (21, 32, 32)   # FlowNetC.py:31
    assert built.correlation_out_shape(8, 8, 4, 1, 4, 1, 1) == (9, 8, 8)          # PWCNet.py:69
    assert built.correlation_out_shape(12, 12, 4, 1, 4, 2, 2) == (5, 6, 6)
    assert built.correlation_out_shape(9, 10, 3, 3, 2, 1, 1) == (5, 9, 10)
    with pytest.raises(ValueError):
        built.correlation_out_shape(4, 4, 0, 1, 4, 1, 1)  # empty output
    # null pointers / bad dims are rejected before any launch
    assert lib.ofr_chnorm_fwd(None, None, 1, 1, 1, 1, None) == 1
    assert lib.ofr_corr_fwd(None, None, None, 1, 1, 8, 8, 4, 1, 4, 1, 1, None, 0, None) == 1
    assert lib.ofr_corr_fwd(1, 1, 1, 1, 1, 8, 8, 4, 2, 4, 1, 1, None, 0, None) == 1  # even kernel size
    assert lib.ofr_corr_workspace(24, 32, 128, 128, 4, 1, 4, 1, 1, 0) == 0        # PWC path needs no scratch
    assert lib.ofr_corr_workspace(1, 4, 32, 32, 20, 1, 20, 1, 2, 0) >= 4 * (2 * 4 + 441) * 32 * 32
    # FlowNetC shapes that tile (class images multiples of 8 x 16, C % 16 == 0) run on the tensor-core kernels, which
    # address every tensor in place: no scratch either; C > 256 only in the forward
    assert lib.ofr_corr_workspace(24, 256, 64, 64, 20, 1, 20, 1, 2, 0) == 0
    assert lib.ofr_corr_workspace(24, 256, 64, 64, 20, 1, 20, 1, 2, 1) == 0
    assert lib.ofr_corr_workspace(1, 512, 32, 32, 20, 1, 20, 1, 2, 0) == 0
    assert lib.ofr_corr_workspace(1, 512, 32, 32, 20, 1, 20, 1, 2, 1) > 0
    assert lib.ofr_corr_workspace(1, 256, 24, 24, 20, 1, 20, 1, 2, 0) > 0           # 12 x 12 class images do not tile
    assert lib.ofr_resample_fwd(None, None, None, 1, 1, 4, 4, 4, 4, 0, None) == 1
    assert lib.ofr_loss_workspace(0) == 0 and lib.ofr_loss_workspace(9) == 0
    assert lib.ofr_loss_workspace(7) > 0
    assert lib.ofr_resample_bwd_workspace(24, 32, 128, 128, 128, 128, 2) > 24 * 128 * 128 * 4 * 3


def test_workspace_matches_oracle_shape_rule(built):
    import oracle
    for args in [(32, 32, 20, 1, 20, 1, 2), (64, 48, 4, 1, 4, 1, 1), (17, 23, 3, 3, 2, 2, 1), (30, 30, 0, 1, 2, 3, 1)]:
        assert built.correlation_out_shape(*args) == oracle.correlation_out_shape(*args)


def test_layers_refuse_cpu_tensors(built):
    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        built.ChannelNorm()(x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        built.Correlation(4, 1, 4, 1, 1, 1)(x, x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        built.Resample2d()(x, x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        built.OFEloss([x], [x[:, :1]], x[:, :1])
    with pytest.raises(NotImplementedError):
        built.Resample2d(kernel_size=3)


def test_modules_are_parameter_free(built):
    for m in (built.Correlation(20, 1, 20, 1, 2, 1), built.Resample2d(), built.ChannelNorm()):
        assert len(m.state_dict()) == 0 and len(list(m.parameters())) == 0
    c = built.Correlation()  # upstream defaults
    assert (c.pad_size, c.kernel_size, c.max_displacement, c.stride1, c.stride2, c.corr_multiply) == (0, 0, 0, 1, 2, 1)


def test_product_never_imports_oracle(built):
    pkg_dir = os.path.dirname(built.LIB_PATH)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f"{f} imports the oracle"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_reference_models_import_with_shims(built):
    """The reference's own model files construct with ofr_b200 layers at the import paths they expect."""
    import types
    built.install_shims()
    added = [REF, os.path.join(REF, "PWC", "models")]
    sys.path[:0] = added
    try:
        import PWCNet
        net = PWCNet.PWCDCNet(md=4)
        assert isinstance(net.corr, built.Correlation)
        assert sum(p.numel() for p in net.parameters()) == 9374184
        import flownet2.models as fm
        net2 = fm.FlowNet2(types.SimpleNamespace(fp16=False, rgb_max=255.0), batchNorm=True)
        assert isinstance(net2.channelnorm, built.ChannelNorm)
        assert isinstance(net2.resample1, built.Resample2d)
        assert isinstance(net2.flownetc.corr, built.Correlation)
        assert all("corr" not in k and "resample" not in k and "channelnorm" not in k for k in net2.state_dict())
    finally:
        for p in added:
            sys.path.remove(p)
        for name in ["PWCNet"] + [n for n in list(sys.modules) if n == "flownet2" or n.startswith("flownet2.")]:
            sys.modules.pop(name, None)
        built.uninstall_shims()

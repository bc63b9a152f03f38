"""Import the reference's own model files, unmodified, with the ofr_b200 layers at the import paths they expect.

Measurement / test harness only.  The files come from ``baseline/_ref/`` (staged byte for byte by
``tools/stage_reference.py``; git-ignored, travels to the GPU box) or, in the dev container, from
``/root/reference``.  Nothing in the reference files is edited; what is done around them:

* ``ofr_b200.install_shims()`` registers Correlation / Resample2d / ChannelNorm at the module paths the reference
  imports (they do not exist in its tree);
* third-party imports that are absent in this image and irrelevant to the hot path (monai, matplotlib, skimage,
  torchmetrics, cv2, RAFT) are satisfied by empty stand-in modules, so ``import models`` / ``import loss`` succeed.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(_HERE))
STAGED = os.path.join(REPO_ROOT, "baseline", "_ref")
_ABSENT = ["monai", "monai.data", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "skimage", "skimage.measure",
           "skimage.transform", "torchmetrics", "torchmetrics.functional", "torchmetrics.functional.clustering",
           "torchmetrics.functional.regression", "RAFT", "RAFT.core", "RAFT.core.raft", "cv2"]
_REF_TOP = ["PWCNet", "PWC", "flownet2", "FlowNetS", "models", "loss", "utils", "networks"]


def reference_dir():
    """Directory holding the reference's files, or None (then the callers skip)."""
    for cand in (STAGED, "/root/reference"):
        if os.path.exists(os.path.join(cand, "PWC", "models", "PWCNet.py")):
            return cand
    return None


class _Sink:
    """Attribute sink standing in for anything imported from an absent third-party module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Sink()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Sink()


def _module_getattr(k):
    if k.startswith("__"):
        raise AttributeError(k)
    return _Sink()


def _stub_absent():
    import torch
    made = []
    for name in _ABSENT:
        if name in sys.modules:
            continue
        try:
            if "." not in name:
                importlib.import_module(name)
                continue
        except ImportError:
            pass
        m = types.ModuleType(name)
        m.__path__ = []
        m.__ofr_b200_refstub__ = True
        m.__getattr__ = _module_getattr
        sys.modules[name] = m
        made.append(name)
    if "monai.data" in made:
        sys.modules["monai.data"].MetaTensor = torch.Tensor
    return made


class ReferenceEnv:
    """Context manager: inside it ``import PWCNet``, ``import flownet2.models``, ``import loss``, ``import models``
    resolve to the reference's files with the ofr_b200 layers shimmed in; on exit every trace is removed."""

    def __init__(self, layers="ofr"):
        """``layers``: "ofr" = the CUDA layers (product), or a dict {"Correlation": cls, "Resample2d": cls,
        "ChannelNorm": cls, "spatial_correlation_sample": fn} of stand-ins (the tests pass the oracle's)."""
        self.layers = layers
        self.ref = reference_dir()

    def __enter__(self):
        if self.ref is None:
            raise RuntimeError("reference files not found: run `python tools/stage_reference.py` in the dev container")
        from .. import shims
        self._paths = [self.ref, os.path.join(self.ref, "PWC", "models")]
        sys.path[:0] = self._paths
        self._stubs = _stub_absent()
        if self.layers == "ofr":
            shims.install()
        else:
            saved = {k: dict(v) for k, v in shims._LEAVES.items()}
            try:
                shims._LEAVES["correlation_package.correlation"]["Correlation"] = self.layers["Correlation"]
                shims._LEAVES["resample2d_package.resample2d"]["Resample2d"] = self.layers["Resample2d"]
                shims._LEAVES["channelnorm_package.channelnorm"]["ChannelNorm"] = self.layers["ChannelNorm"]
                shims.install()
            finally:
                for k, v in saved.items():
                    shims._LEAVES[k].clear()
                    shims._LEAVES[k].update(v)
        return self

    def __exit__(self, *exc):
        from .. import shims
        for p in self._paths:
            if p in sys.path:
                sys.path.remove(p)
        for name in list(sys.modules):
            top = name.split(".")[0]
            if top in _REF_TOP or getattr(sys.modules[name], "__ofr_b200_refstub__", False):
                f = getattr(sys.modules[name], "__file__", None) or ""
                if getattr(sys.modules[name], "__ofr_b200_refstub__", False) or f.startswith(self.ref) or \
                        getattr(sys.modules[name], "__ofr_b200_shim__", False):
                    del sys.modules[name]
        shims.uninstall()
        return False

    # -- constructors of the reference's predictors, exactly as models.py:205-252 builds them (minus the checkpoint
    # loads from C:\Users\... paths)
    def pwcnet(self):
        import PWCNet  # PWC/models/PWCNet.py, found through the sys.path entry models.py:9 appends
        return PWCNet.PWCDCNet(md=4)

    def flownet2(self):
        import flownet2.models as fm
        return fm.FlowNet2(types.SimpleNamespace(fp16=False, rgb_max=255.0), batchNorm=True)

    def loss_module(self):
        return importlib.import_module("loss")

    def models_module(self):
        return importlib.import_module("models")

"""Import-path shims: make the reference's model files load these layers unchanged.

The reference imports the custom layers from module paths that do not exist in its tree
(SURVEY.md section 0):

    flownet2/networks/FlowNetC.py:8   from .correlation_package.correlation import Correlation
    flownet2/models.py:10-11,20-21    from [.]networks.resample2d_package.resample2d import Resample2d
                                      from [.]networks.channelnorm_package.channelnorm import ChannelNorm
    PWC/models/PWCNet.py:13           from correlation_package.correlation import Correlation
    FlowNetS/util.py:5                from spatial_correlation_sampler import spatial_correlation_sample

``install()`` registers modules under exactly those names in ``sys.modules`` (the reference tree
is read-only, so nothing is written into it).  Call it before importing the reference's models.
"""
import importlib.machinery
import sys
import types

from . import layers

_LEAVES = {
    "correlation_package.correlation": {"Correlation": layers.Correlation},
    "resample2d_package.resample2d": {"Resample2d": layers.Resample2d},
    "channelnorm_package.channelnorm": {"ChannelNorm": layers.ChannelNorm},
}
# package prefixes under which the reference looks the three packages up
_PREFIXES = ["", "networks.", "flownet2.networks."]


def _module(name, attrs=None, package=False):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=package)
        if package:
            m.__path__ = []
        m.__ofr_b200_shim__ = True
        sys.modules[name] = m
    if attrs:
        m.__dict__.update(attrs)
    return m


def install(prefixes=None):
    """Register the shim modules; returns the list of module names now served by ofr_b200."""
    served = []
    for prefix in (prefixes if prefixes is not None else _PREFIXES):
        for leaf, attrs in _LEAVES.items():
            pkg_name, mod_name = leaf.split(".")
            full_pkg = prefix + pkg_name
            full_mod = prefix + leaf
            # never create a fake parent package such as "flownet2.networks": the real one must import
            pkg = _module(full_pkg, package=True)
            mod = _module(full_mod, attrs)
            setattr(pkg, mod_name, mod)
            parent = sys.modules.get(prefix.rstrip(".")) if prefix else None
            if parent is not None:
                setattr(parent, pkg_name, pkg)
            served += [full_pkg, full_mod]
    _module("spatial_correlation_sampler", {"spatial_correlation_sample": layers.spatial_correlation_sample})
    served.append("spatial_correlation_sampler")
    return served


def uninstall():
    for name in [n for n, m in sys.modules.items() if getattr(m, "__ofr_b200_shim__", False)]:
        del sys.modules[name]

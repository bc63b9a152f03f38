"""B200-native (sm_100a) hot path of self-supervised brain-MRI registration with deep optical flow.

Only what the hot path needs: the three flownet2 custom layers, the two
in-tree warps and the self-supervised loss, each a thin Python module over hand-written CUDA kernels
reached through the C ABI of ``include/ofr_b200.h``.

    from ofr_b200 import Correlation, Resample2d, ChannelNorm, stn_warp, pwc_warp, OFEloss
    ofr_b200.install_shims()     # then `import flownet2.models` / `PWCNet` from the reference work unchanged
"""
from . import _lib
from ._lib import build, launch_count, LIB_PATH, EXPORTS
from .layers import (Correlation, Resample2d, ChannelNorm, stn_warp, stn_warp_multi, pwc_warp, channelnorm_diff, spatial_correlation_sample,
                     correlation_out_shape)
from .loss import OFEloss, registration_loss, photometric_loss, correlation_loss, smoothness_loss
from .registration import OpticalFlowReg, grid_generator, round_clamp_labels
from .pairs import ElasticPairGenerator, control_grid_size
from .flowhead import FlowConv2d, FlowUpsample2d, accelerate_flow_heads
from .shims import install as install_shims, uninstall as uninstall_shims

__all__ = [
    "Correlation", "Resample2d", "ChannelNorm", "stn_warp", "stn_warp_multi", "pwc_warp", "channelnorm_diff", "spatial_correlation_sample",
    "correlation_out_shape", "OFEloss", "registration_loss", "photometric_loss", "correlation_loss", "smoothness_loss",
    "OpticalFlowReg", "grid_generator", "round_clamp_labels", "ElasticPairGenerator", "control_grid_size", "FlowConv2d", "FlowUpsample2d", "accelerate_flow_heads",
    "install_shims", "uninstall_shims", "build", "launch_count", "LIB_PATH", "EXPORTS",
]

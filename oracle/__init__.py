"""CPU oracle for the optical-flow registration hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU in plain PyTorch / numpy, the arithmetic of the
layers the reference calls on its hot path.  It exists so
the CUDA path can be *checked*; it is never the thing shipped or measured.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under the
product package imports it, and the product raises when its CUDA library is
missing rather than falling back to anything in here.

Parity pinning status (see DESIGN.md "Oracle"):

* ``warp_ref.stn_warp``, ``warp_ref.pwc_warp`` and everything in ``loss_ref``
  follow code that IS in the reference tree (``models.py:195-204,256-268``,
  ``PWC/models/PWCNet.py:143-179``, ``loss.py:9-84``).  They are PINNED: the
  fixtures under ``tests/golden/`` were produced by importing and running those
  reference functions unmodified (``tests/golden/make_golden.py``).
* ``layers_ref.correlation``, ``layers_ref.resample2d`` and
  ``layers_ref.channelnorm`` restate layers the reference only *imports* from
  the un-vendored, unpinned ``NVIDIA/flownet2-pytorch`` (``README.md:18-19``;
  import sites ``flownet2/networks/FlowNetC.py:8``, ``flownet2/models.py:10-11``,
  ``PWC/models/PWCNet.py:13``).  No source, test or golden vector for them
  exists under ``/root/reference``: **parity unpinned** against upstream.  They
  are anchored instead on (i) an independent numpy statement of the Caffe
  correlation operator that ships in this image (TVM ``correlation_nchw_python``)
  and (ii) ``F.grid_sample(padding_mode='border', align_corners=True)`` for the
  clamped bilinear gather; fixtures from both are committed under
  ``tests/golden/``.
* ``flowhead_ref.predict_flow`` / ``deconv`` ARE the reference's ops: it builds these layers from
  ``torch.nn`` (``PWC/models/PWCNet.py:30-34``).  PINNED by construction; the loop restatements beside them fix the
  index conventions the CUDA kernels implement.
"""
from .layers_ref import (correlation, correlation_loops, resample2d, resample2d_loops,
                         channelnorm, channelnorm_backward, correlation_out_shape, MODULES, Correlation_like)
from .warp_ref import stn_warp, pwc_warp, pixel_grid
from .loss_ref import (charbonnier, photometric, smoothness, ncc, ofe_loss)
from .flowhead_ref import predict_flow, deconv, predict_flow_loops, deconv_loops

__all__ = [
    "correlation", "correlation_loops", "correlation_out_shape", "resample2d", "resample2d_loops",
    "channelnorm", "channelnorm_backward", "stn_warp", "pwc_warp", "pixel_grid",
    "charbonnier", "photometric", "smoothness", "ncc", "ofe_loss", "MODULES", "Correlation_like",
    "predict_flow", "deconv", "predict_flow_loops", "deconv_loops",
]

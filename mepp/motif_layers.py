"""Same interface as the reference's mepp/motif_layers.py, implemented by mepp_b200.motif_layers."""
from mepp_b200.motif_layers import *  # noqa: F401,F403
from mepp_b200 import motif_layers as _impl

__doc__ = _impl.__doc__ or __doc__

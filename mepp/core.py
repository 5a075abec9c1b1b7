"""Same interface as the reference's mepp/core.py, implemented by mepp_b200.core."""
from mepp_b200.core import *  # noqa: F401,F403
from mepp_b200 import core as _impl

__doc__ = _impl.__doc__ or __doc__

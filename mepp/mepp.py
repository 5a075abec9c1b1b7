"""Same interface as the reference's mepp/mepp.py, implemented by mepp_b200.mepp."""
from mepp_b200.mepp import *  # noqa: F401,F403
from mepp_b200 import mepp as _impl

__doc__ = _impl.__doc__ or __doc__

"""Same interface as the reference's mepp/permutations.py, implemented by mepp_b200.permutations."""
from mepp_b200.permutations import *  # noqa: F401,F403
from mepp_b200 import permutations as _impl

__doc__ = _impl.__doc__ or __doc__

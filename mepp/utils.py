"""Same interface as the reference's mepp/utils.py, implemented by mepp_b200.utils."""
from mepp_b200.utils import *  # noqa: F401,F403
from mepp_b200 import utils as _impl

__doc__ = _impl.__doc__ or __doc__

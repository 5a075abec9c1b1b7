"""Same interface as the reference's mepp/io.py, implemented by mepp_b200.io."""
from mepp_b200.io import *  # noqa: F401,F403
from mepp_b200 import io as _impl

__doc__ = _impl.__doc__ or __doc__

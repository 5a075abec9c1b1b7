"""Same interface as the reference's mepp/batch.py, implemented by mepp_b200.batch."""
from mepp_b200.batch import *  # noqa: F401,F403
from mepp_b200 import batch as _impl

__doc__ = _impl.__doc__ or __doc__

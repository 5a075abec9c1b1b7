"""Same interface as the reference's mepp/onehot_dna.py, implemented by mepp_b200.onehot_dna."""
from mepp_b200.onehot_dna import *  # noqa: F401,F403
from mepp_b200 import onehot_dna as _impl

__doc__ = _impl.__doc__ or __doc__

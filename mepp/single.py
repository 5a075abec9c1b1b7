"""Same interface as the reference's mepp/single.py, implemented by mepp_b200.single."""
from mepp_b200.single import *  # noqa: F401,F403
from mepp_b200 import single as _impl

__doc__ = _impl.__doc__ or __doc__

if __name__ == "__main__":
    import sys
    sys.exit(_impl._main())  # `python -m mepp.single -y params.yaml` (single.py:304-306, 410-412)

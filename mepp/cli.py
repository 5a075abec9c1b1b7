"""Same interface as the reference's mepp/cli.py, implemented by mepp_b200.cli."""
from mepp_b200.cli import *  # noqa: F401,F403
from mepp_b200 import cli as _impl

__doc__ = _impl.__doc__ or __doc__

if __name__ == "__main__":
    import sys
    sys.exit(_impl.main())  # pragma: no cover

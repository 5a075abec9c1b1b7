import sys

from mepp_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())  # pragma: no cover

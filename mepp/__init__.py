"""Drop-in import name for the reference package: `import mepp`, `from mepp.core import ...`, `python -m mepp`
(npdeloss/mepp setup.py:51-52, mepp/__init__.py).  Every module here re-exports the module of the same name from
mepp_b200, the B200-native implementation of the profiling path; the reference's plotting / HTML / clustering /
motif-learning modules are outside that path and are not provided."""
__author__ = "mepp_b200"
__version__ = "0.0.1"

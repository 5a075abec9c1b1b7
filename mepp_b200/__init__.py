"""mepp_b200 -- B200-native engine for the MEPP profiling path (drop-in for npdeloss/mepp's
mepp/core.py, batch.py, single.py entry points).  Hand-written sm_100a CUDA behind a C ABI
(include/mepp_b200.h); no CPU fallback."""
__version__ = "0.1.0"

"""CPU oracle for the MEPP profiling hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's profiling
path (npdeloss/mepp: mepp/core.py:490-632, mepp/motif_layers.py:12-165,
mepp/utils.py:53-308, mepp/permutations.py:5-58, mepp/io.py:19-95) and of the
two MOODS-python functions it calls (MOODS.tools.log_odds / threshold_from_p,
MOODS-python >= 1.9.4, not vendored by the reference).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product (`mepp_b200/`) never does.

PARITY UNPINNED: the reference ships no golden vectors or numeric tests for
this path (tests/test_mepp.py:14-33 is a stub) and cannot be imported in this
image (TensorFlow, MOODS, Biopython absent, no network).  The oracle is pinned
only by hand-derivable known-answer tests (tests/test_oracle_kat.py) and by a
brute-force re-derivation of the MOODS threshold DP.
"""

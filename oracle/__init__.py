"""CPU oracle for the MEPP profiling hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's profiling
path (npdeloss/mepp: mepp/core.py:490-632, mepp/motif_layers.py:12-165,
mepp/utils.py:53-308, mepp/permutations.py:5-58, mepp/io.py:19-95) and of the
two MOODS-python functions it calls (MOODS.tools.log_odds / threshold_from_p,
MOODS-python >= 1.9.4, not vendored by the reference).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product (`mepp_b200/`) never does.

PARITY PARTLY PINNED.  The reference ships no golden vectors or numeric tests for this path
(tests/test_mepp.py:14-33 is a stub) and its package cannot be imported in this image (TensorFlow, MOODS,
Biopython absent, no network).  But the stages of mepp/core.py that are plain numpy / pandas / scipy were
executed here, unmodified, by tests/golden/make_reference_stat_fixtures.py (TensorFlow's zeros / reduce_sum /
sqrt mapped to numpy) and their outputs are committed as tests/golden/reference_stats.npz:
  pinned   -- positional / partial correlation (core.py:182-267), permutation statistics (core.py:269-376),
              parametric interval and t-test (core.py:380-413), counts, density, smoothing, ranks
              (core.py:172-180, 415-488), N rule / one-hot / degenerate filter / ordering (onehot_dna.py:21-70,
              utils.py:53-128): tests/test_reference_golden.py (oracle, host and device paths);
  UNPINNED -- the scan stage: Keras Conv1D model, MOODS log-odds and threshold (motif_layers.py:12-165), the
              Biopython parsers.  These rest on hand-derivable known-answer tests (tests/test_oracle_kat.py) and a
              brute-force re-derivation of the MOODS threshold DP.
"""

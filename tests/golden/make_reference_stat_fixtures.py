"""Golden vectors from the REFERENCE'S OWN CODE for the stages of the path that are plain numpy / pandas / scipy.

Run in the build container (needs /root/reference; nothing at test time does):
    python tests/golden/make_reference_stat_fixtures.py
writes tests/golden/reference_stats.npz.

mepp/core.py is loaded from /root/reference unmodified.  Its module-level imports of TensorFlow and of the
TF / MOODS based sibling modules are satisfied by stand-ins: `tensorflow.zeros`, `tensorflow.math.reduce_sum`
and `tensorflow.math.sqrt` map to the numpy functions of the same meaning (float32 in, float32 out; `+=` on the
zeros rebinds instead of adding in place, like on an immutable TF tensor), which is all
`get_positional_correlation_from_processed_dataset` (core.py:195-267) uses; everything else the fixtures call
(`get_positional_r_df` core.py:269-376, `add_parametric_confidence_intervals_to_positional_r_df` core.py:380-413,
counts / density / smoothing core.py:415-488, `get_score_rank_df` core.py:172-180) is numpy / pandas / scipy only.
Shim: `np.float` (removed from numpy >= 1.24, used at core.py:424,464) is set to `float`.
What this does NOT cover: the Keras conv model and the MOODS threshold (TensorFlow and MOODS are not installable
here) -- the scan stage of the oracle stays unpinned.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/mepp/core.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_stats.npz")


def load_reference_core():
    np.float = float                                                  # core.py:424,464
    class Tensor(np.ndarray):
        """TensorFlow tensors are immutable: `a += b` rebinds a to the (broadcast) sum (core.py:238-253 relies on it)."""
        def __iadd__(self, other):
            return np.add(self, other).view(Tensor)

    tf = types.ModuleType("tensorflow")
    tf.zeros = lambda shape: np.zeros(tuple(shape), dtype=np.float32).view(Tensor)
    tfm = types.ModuleType("tensorflow.math")
    tfm.reduce_sum = lambda x, axis=None: np.sum(x, axis=axis)
    tfm.sqrt = np.sqrt
    tfk = types.ModuleType("tensorflow.keras")
    tfl = types.ModuleType("tensorflow.keras.layers")
    tfl.AveragePooling1D = object
    tf.math, tf.keras, tfk.layers = tfm, tfk, tfl
    pkg = types.ModuleType("mepp")
    pkg.__path__ = [os.path.dirname(REF)]
    ml = types.ModuleType("mepp.motif_layers")
    ml.motif_matrix_to_conv_model = None
    ut = types.ModuleType("mepp.utils")
    ut.get_sequence_length_from_dataset = None
    sys.modules.update({"tensorflow": tf, "tensorflow.math": tfm, "tensorflow.keras": tfk,
                        "tensorflow.keras.layers": tfl, "mepp": pkg, "mepp.motif_layers": ml, "mepp.utils": ut})
    spec = importlib.util.spec_from_file_location("mepp.core", REF)
    core = importlib.util.module_from_spec(spec)
    sys.modules["mepp.core"] = core
    spec.loader.exec_module(core)
    return core


def load_reference_staging():
    """mepp/onehot_dna.py as is (pure numpy) and mepp/utils.py with TensorFlow / Keras / statsmodels names stubbed
    (only its pandas functions are called: utils.py:53-128, 391-446)."""
    class Anything(types.ModuleType):
        def __getattr__(self, name):
            return None
    del sys.modules["mepp.utils"]
    sys.modules["tensorflow.keras"] = Anything("tensorflow.keras")
    for name in ("statsmodels", "statsmodels.stats", "statsmodels.stats.multitest"):
        sys.modules[name] = Anything(name)
    mods = {}
    for name in ("onehot_dna", "utils"):
        spec = importlib.util.spec_from_file_location(f"mepp.{name}", os.path.join(os.path.dirname(REF), f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"mepp.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["onehot_dna"], mods["utils"]


class FakeDataset(list):
    """Iterable of batches with the `.element_spec[i].shape` the reference reads (core.py:201-203)."""
    def __init__(self, batches):
        super().__init__(batches)
        self.element_spec = tuple(types.SimpleNamespace(shape=(None,) + tuple(a.shape[1:])) for a in batches[0])


def main():
    core = load_reference_core()
    rng = np.random.default_rng(20260101)
    out = {}
    # ---- permutation / parametric statistics (core.py:269-413)
    cases = [("a", 41, 24, None, 95, 500), ("b", 17, 1, 3, 95, 64), ("c", 30, 100, None, 80, 2000), ("d", 25, 0, None, 95, 300)]
    for name, L, P, center, ci, n in cases:
        r = rng.normal(scale=0.05, size=(L, 1 + P)).astype(np.float32)
        r[0, :] = 0.0                                                  # a position without any match: r = 0 everywhere
        if P == 0:
            r = r[:, 0]
        df = core.get_positional_r_df(r, center=center, confidence_interval_pct=ci)
        df = core.add_parametric_confidence_intervals_to_positional_r_df(df, n, confidence_interval_pct=ci)
        out[f"stat_{name}__r"] = r
        out[f"stat_{name}__args"] = np.array([L, P, -1 if center is None else center, ci, n], dtype=np.int64)
        out[f"stat_{name}__columns"] = np.array(list(df.columns))
        for col in df.columns:
            out[f"stat_{name}__col__{col}"] = df[col].to_numpy()
    # ---- counts / density / smoothing / ranks (core.py:172-180, 415-488)
    X0 = rng.random((60, 33)).astype(np.float32)
    X0[rng.random(X0.shape) > 0.07] = 0.0
    cdf = core.get_motif_counts_by_position_df(X0)
    out["cnt__X0"] = X0
    out["cnt__position"] = cdf["position"].to_numpy()
    out["cnt__count"] = cdf["count"].to_numpy()
    for w in (3, 5, 11):
        out[f"cnt__smoothed_w{w}"] = core.smooth_motif_counts_by_position_df(cdf, window=w)["count"].to_numpy()
    ddf = core.get_motif_density_by_rank_df(X0)
    out["cnt__density"] = ddf["density"].to_numpy()
    out["cnt__density_smoothed_w5"] = core.smooth_motif_density_by_rank_df(ddf, 5)["density"].to_numpy()
    scores = rng.normal(size=60).astype(np.float32)
    sdf = core.get_score_rank_df(scores)
    out["rank__scores"] = scores
    out["rank__rank"] = sdf["rank"].to_numpy()
    # ---- positional / partial correlation (core.py:182-267), two batches, with and without the GC covariate
    N, L, C = 200, 12, 6
    X = rng.random((N, L, 1)).astype(np.float32)
    X[rng.random(X.shape) > 0.2] = 0.0
    Y = rng.normal(loc=3.0, scale=2.0, size=(N, 1, C)).astype(np.float32)
    Z = (0.4 + 0.1 * rng.random((N, 1, 1))).astype(np.float32)
    out["corr__X"], out["corr__Y"], out["corr__Z"] = X, Y, Z
    halves = [slice(0, 120), slice(120, N)]
    out["corr__r_xy_z"] = np.asarray(core.get_positional_correlation_from_processed_dataset(
        FakeDataset([(X[s], Y[s], Z[s]) for s in halves])))
    out["corr__r_xy"] = np.asarray(core.get_positional_correlation_from_processed_dataset(
        FakeDataset([(X[s], Y[s]) for s in halves])))
    # ---- staging: N rule, one-hot, degenerate filter, ordering (onehot_dna.py:21-70, utils.py:53-128)
    onehot_dna, utils = load_reference_staging()
    import pandas as pd
    seqs = ["ACGTACGTAC", "acgtNNRYAC", "ACGTNACGTN", "TTTTTTTTTT", "ACGUACGTAC", "ACG", "GGGGCCCCAA", "NNNNNNNNNN"]
    out["oh__seqs"] = np.array(seqs)
    out["oh__mats"] = np.array([np.array(onehot_dna.seq_to_mat(s_.ljust(10, "N")[:10])) for s_ in seqs])
    out["oh__degenerate_pct"] = np.array([onehot_dna.get_degenerate_pct(s_) for s_ in seqs])
    out["oh__n_mask"] = np.array([onehot_dna.masked_seq_to_n_mask(s_) for s_ in seqs])
    ids = [f"s{i}" for i in range(len(seqs))]
    score = dict(zip(ids, [0.5, -1.25, 3.0, 2.0, 0.25, 9.0, -3.5, 1.0]))
    df = utils.scored_fasta_dicts_to_df(dict(zip(ids, seqs)), score, {i: f"d {i}" for i in ids})
    for thr in (100, 25, 0):
        f = utils.filter_scored_fasta_df(df, degenerate_pct_thresh=thr)
        out[f"df__filtered_ids_thr{thr}"] = f["sequence_id"].to_numpy().astype(str)
        out[f"df__filtered_pct_thr{thr}"] = f["degenerate_pct"].to_numpy()
        o = utils.order_scored_fasta_df(f, shuffle=False)
        out[f"df__ordered_ids_thr{thr}"] = o["sequence_id"].to_numpy().astype(str)
        out[f"df__ordered_rank_thr{thr}"] = o["rank"].to_numpy()
    # ---- per-motif summary of a positions table (utils.py:391-446)
    for name in ("a", "c", "d"):
        cols = [str(c) for c in out[f"stat_{name}__columns"]]
        pdf = pd.DataFrame({c: out[f"stat_{name}__col__{c}"] for c in cols})
        mm = utils.get_minmax_stats(pdf)
        out[f"minmax_{name}__keys"] = np.array(list(mm.keys()))
        out[f"minmax_{name}__values"] = np.array([np.nan if v is None else float(v) for v in mm.values()], dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()

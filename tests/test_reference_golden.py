"""The oracle (and the host / device statistics) against golden vectors produced by the REFERENCE'S OWN CODE
(tests/golden/make_reference_stat_fixtures.py ran mepp/core.py unmodified on these inputs, TensorFlow's three array
functions mapped to numpy).  This pins stages a12-a16 of SURVEY.md section 8: positional / partial correlation,
permutation statistics, parametric interval, counts, density, smoothing.  The scan stage (Keras conv, MOODS threshold)
cannot be pinned here."""
import os

import numpy as np
import pytest

from oracle import profile

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_stats.npz")


@pytest.fixture(scope="module")
def ref():
    with np.load(FIX) as z:
        return {k: z[k] for k in z.files}


CASES = ["a", "b", "c", "d"]


def _case(ref, name):
    L, P, center, ci, n = (int(v) for v in ref[f"stat_{name}__args"])
    cols = {str(c): ref[f"stat_{name}__col__{c}"] for c in ref[f"stat_{name}__columns"]}
    return ref[f"stat_{name}__r"], P, (None if center < 0 else center), ci, n, cols


@pytest.mark.parametrize("name", CASES)
def test_oracle_statistics_equal_the_reference(ref, name):
    r, P, center, ci, n, want = _case(ref, name)
    got = profile.positional_r_columns(r, center=center, confidence_interval_pct=ci)
    lo, up, pv = profile.parametric_ci(got['positional_r'], n, ci)
    got['positional_r_lower'], got['positional_r_upper'], got['positional_r_pval'] = lo, up, pv
    assert list(got.keys()) == list(want.keys())                       # same columns, same order (core.py:286-411)
    for k in want:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        if k in ('positional_r_lower', 'positional_r_upper'):
            # the oracle restates the float32 arithmetic of the reference's pinned numpy (< 1.24); under this
            # container's numpy 2 the same reference lines promote to float64: equal to float32 precision
            assert np.abs(a - b).max() <= 3e-7, k
        elif k == 'positional_r_pval':
            assert np.allclose(a, b, rtol=2e-5, atol=1e-300), k
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("name", CASES)
def test_host_finalisation_equals_the_reference(ref, name):
    """stats_host.finalize_into from order statistics / counts computed with numpy sorts (what K3 delivers)."""
    from mepp_b200 import stats_host
    r, P, center, ci, n, want = _case(ref, name)
    r2 = r if r.ndim > 1 else r[:, None]
    L = r2.shape[0]
    store = stats_host.alloc_positions(1, L, P)
    count = np.arange(L, dtype=np.int32)[None] % 5
    args = [None] * 5
    if P > 0:
        q = stats_host.percentile_q(ci)
        k_lo, _ = stats_host.virtual_index(P, q[0]); k_hi, _ = stats_host.virtual_index(P, q[1])
        ks_lo, _ = stats_host.virtual_index(L * P, q[0]); ks_hi, _ = stats_host.virtual_index(L * P, q[1])
        srt = np.sort(r2[:, 1:], axis=-1)
        pooled = np.sort(r2[:, 1:].ravel())
        M = L * P
        full_os = np.stack([srt[:, k_lo], srt[:, min(k_lo + 1, P - 1)], srt[:, k_hi], srt[:, min(k_hi + 1, P - 1)]], -1)[None]
        semi_os = np.array([[pooled[ks_lo], pooled[min(ks_lo + 1, M - 1)], pooled[ks_hi], pooled[min(ks_hi + 1, M - 1)]]])
        full_cnt = (np.abs(r2[:, 1:]) >= np.abs(r2[:, :1])).sum(-1)[None]
        semi_cnt = (np.abs(r2[:, 1:]).ravel()[None, :] >= np.abs(r2[:, :1])).sum(-1)[None]
        a64 = np.abs(r2.astype(np.float64))
        integral = (0.5 * (a64[1:] + a64[:-1])).sum(0)[None]
        args = [full_os.astype(np.float32), full_cnt, semi_os.astype(np.float32), semi_cnt, integral]
    stats_host.finalize_into(store, 0, r2[:, 0][None].copy(), *args, count, n, 0, ci)
    got = stats_host.positions_view(store, center)
    for k in want:
        a, b = np.asarray(got[k][0], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        if k in ('positional_r_lower', 'positional_r_upper'):
            assert np.abs(a - b).max() <= 3e-7, k
        elif k == 'positional_r_pval':
            assert np.allclose(a, b, rtol=2e-5, atol=1e-300), k
        elif k in ('integral_r', 'permutation_integral_r_lower', 'permutation_integral_r_upper'):
            assert np.allclose(a, b, rtol=3e-7), k                    # float32 trapezoid sums: summation order
        else:
            assert np.array_equal(a, b), k


def test_oracle_counts_density_smoothing_equal_the_reference(ref):
    X0 = ref["cnt__X0"]
    cnt = profile.counts_by_position(X0)
    assert np.array_equal(cnt, ref["cnt__count"])
    assert np.array_equal(np.arange(X0.shape[1]) - X0.shape[1] // 2, ref["cnt__position"])
    for w in (3, 5, 11):
        assert np.allclose(profile.smooth_counts(cnt, w), ref[f"cnt__smoothed_w{w}"], rtol=1e-13, atol=1e-13), w
    assert np.array_equal(profile.density_by_rank(X0), ref["cnt__density"])
    from mepp_b200 import stats_host
    assert np.allclose(stats_host.smooth_counts(cnt[None], 5)[0], ref["cnt__smoothed_w5"], rtol=1e-13, atol=1e-13)
    assert np.array_equal(np.argsort(ref["rank__scores"]), ref["rank__rank"])          # core.py:172-180


def test_oracle_correlation_equals_the_reference(ref):
    """core.py:182-267 evaluated by the reference in float32 raw moments over two batches; the oracle evaluates
    the same formulas in float64.  They agree to the float32 raw-moment noise of the reference."""
    X, Y, Z = ref["corr__X"][:, :, 0], ref["corr__Y"][:, 0, :], ref["corr__Z"][:, 0, 0]
    got_z = profile.positional_correlation(X, Y, Z)
    got = profile.positional_correlation(X, Y, None)
    assert np.abs(got_z - ref["corr__r_xy_z"]).max() <= 5e-6          # measured: 6.3e-7 at |r| <= 0.18
    assert np.abs(got - ref["corr__r_xy"]).max() <= 5e-6            # measured: 1.9e-7
    assert np.abs(ref["corr__r_xy"]).max() > 0.05                     # (a non-trivial signal)


def test_staging_rules_equal_the_reference(ref):
    """N rule / one-hot (onehot_dna.py:21-70), degenerate filter and ordering (utils.py:53-128): the reference's own
    functions against the ports in mepp_b200 and the oracle's symbol coding."""
    import pandas as pd
    from mepp_b200 import onehot_dna, utils
    from oracle import staging
    seqs = [str(s) for s in ref["oh__seqs"]]
    for s_, mat, pct, nm in zip(seqs, ref["oh__mats"], ref["oh__degenerate_pct"], ref["oh__n_mask"]):
        padded = s_.ljust(10, "N")[:10]
        assert np.array_equal(np.array(onehot_dna.seq_to_mat(padded)), mat)
        assert onehot_dna.get_degenerate_pct(s_) == pct
        assert onehot_dna.masked_seq_to_n_mask(s_) == str(nm)
        # oracle symbols 0..3 = the one-hot channel, 4 = the reference's all-ones "N" row (normalised to 0.25 each)
        codes = staging.ascii_to_codes(np.frombuffer(padded.encode(), dtype=np.uint8)[None])[0]
        want = np.where(mat.sum(1) == 4, 4, mat.argmax(1))
        assert np.array_equal(codes, want)
    ids = [f"s{i}" for i in range(len(seqs))]
    score = dict(zip(ids, [0.5, -1.25, 3.0, 2.0, 0.25, 9.0, -3.5, 1.0]))
    df = utils.scored_fasta_dicts_to_df(dict(zip(ids, seqs)), score, {i: f"d {i}" for i in ids})
    for thr in (100, 25, 0):
        f = utils.filter_scored_fasta_df(df, degenerate_pct_thresh=thr)
        assert list(f["sequence_id"]) == [str(x) for x in ref[f"df__filtered_ids_thr{thr}"]]
        assert np.array_equal(f["degenerate_pct"].to_numpy(), ref[f"df__filtered_pct_thr{thr}"])
        o = utils.order_scored_fasta_df(f, shuffle=False)
        assert list(o["sequence_id"]) == [str(x) for x in ref[f"df__ordered_ids_thr{thr}"]]
        assert np.array_equal(o["rank"].to_numpy(), ref[f"df__ordered_rank_thr{thr}"])


@pytest.mark.parametrize("name", ["a", "c", "d"])
def test_minmax_summary_equals_the_reference(ref, name):
    """summary.get_minmax_stats against utils.get_minmax_stats (utils.py:391-446) on the reference's own table."""
    import pandas as pd
    from mepp_b200 import summary
    cols = [str(c) for c in ref[f"stat_{name}__columns"]]
    pdf = pd.DataFrame({c: ref[f"stat_{name}__col__{c}"] for c in cols})
    got = summary.get_minmax_stats(pdf)
    keys = [str(k) for k in ref[f"minmax_{name}__keys"]]
    assert list(got.keys()) == keys
    for k, v in zip(keys, ref[f"minmax_{name}__values"]):
        g = np.nan if got[k] is None else float(got[k])
        assert (np.isnan(g) and np.isnan(v)) or g == v, k

"""Known-answer tests that pin the CPU oracle (the reference ships no
golden vectors for this path -- PARITY UNPINNED -- so the oracle is pinned by hand-derivable
answers and by an independent brute-force re-derivation of the MOODS threshold DP)."""
import numpy as np
import pytest

from oracle import moods, profile, staging

BG = moods.flat_bg(4)


def consensus_pwm(cons="ACGTACGT"):
    pwm = np.zeros((4, len(cons)))
    for i, c in enumerate(cons):
        pwm["ACGT".index(c), i] = 1.0
    return pwm


def codes_of(seq):
    return staging.ascii_to_codes(np.frombuffer(seq.encode(), dtype=np.uint8)[None, :])


# The DP grid of MOODS' threshold_from_p: 1/1000 (PVAL_DP_MULTIPLIER of MOODS <= 1.9.2, the grid SURVEY.md 8c restates)
# and 1/2000 (DEFAULT_DP_PRECISION of MOODS >= 1.9.3, what the reference's `MOODS-python>=1.9.4` pin most likely runs
# and therefore the default here); every known answer below is derived by hand for both.
GRIDS = [1000.0, 2000.0]


def test_default_grid_is_the_one_of_the_pinned_moods():
    from mepp_b200 import motif_layers
    assert moods.DP_PRECISION == 2000.0 and motif_layers.DP_PRECISION == 2000.0


@pytest.mark.parametrize("grid", GRIDS)
def test_kat1_uniform_pwm_never_matches(grid):
    W = moods.log_odds(np.full((4, 8), 0.25), BG, 1e-4)
    assert np.allclose(W, 0.0, atol=1e-15)
    assert moods.threshold_from_p(W, BG, 1e-4, grid) == pytest.approx(1.0 / grid)
    codes = np.random.default_rng(0).integers(0, 4, (5, 50)).astype(np.int8)
    X0 = profile.scan(codes, W.astype(np.float32), None, 1.0 / grid)
    assert (X0 == 0).all()


@pytest.mark.parametrize("grid", GRIDS)
def test_kat2_consensus_threshold_and_position(grid):
    pwm = consensus_pwm()
    W = moods.log_odds(pwm, BG, 1e-4)
    assert W[0, 0] == pytest.approx(np.log(0.999925 / 0.25) - 0.0, rel=1e-6)  # (1 + 2.5e-5)/1.0001 / .25
    assert W[1, 0] == pytest.approx(-9.2104, abs=1e-3)
    thr = moods.threshold_from_p(W, BG, 1e-4, grid)
    # integer scores: trunc(grid * 1.38622 + 0.5) = 1386 / 2772, trunc(grid * -9.2104 - 0.5) = -9210 / -18421
    i_match, i_mis = {1000.0: (1386, -9210), 2000.0: (2772, -18421)}[grid]
    assert thr == pytest.approx((7 * i_match + i_mis + 1) / grid)      # only exact consensus passes
    assert thr == moods.threshold_bruteforce(W, BG, 1e-4, grid)
    seq = "T" * 10 + "ACGTACGT" + "T" * 12
    X0 = profile.scan(codes_of(seq), W.astype(np.float32), None, thr)
    nz = np.nonzero(X0[0])[0]
    assert list(nz) == [10 + 3]                                   # i + (m-1)//2
    w32 = W.astype(np.float32)
    acc = np.float32(0)
    for j, c in enumerate("ACGTACGT"):
        acc = np.float32(acc + w32["ACGT".index(c), j])
    assert X0[0, 13] == np.float32(acc + np.float32(-thr))


def test_kat3_orientations():
    pwm = consensus_pwm("AACCGGTA")
    rc = "TACCGGTT"                                                # reverse complement of the consensus
    seq = "G" * 7 + rc + "G" * 9
    c = codes_of(seq)
    Wf, Wr, thr = profile.motif_weights(pwm, '+')
    assert Wr is None and (profile.scan(c, Wf, None, thr) == 0).all()
    Wm, _, thr_m = profile.motif_weights(pwm, '-')
    Xm = profile.scan(c, Wm, None, thr_m)
    assert list(np.nonzero(Xm[0])[0]) == [7 + 3]                   # same left pad as '+'
    Wf, Wr, thr_d = profile.motif_weights(pwm, '+/-')
    Xd = profile.scan(c, Wf, Wr, thr_d)
    assert np.array_equal(Xd, Xm) and thr_d == thr


def test_kat4_all_n_sequence():
    pwm = consensus_pwm()
    Wf, _, thr = profile.motif_weights(pwm, '+')
    c = codes_of("N" * 30)
    assert (c == 4).all()
    assert (profile.scan(c, Wf, None, thr) == 0).all()
    assert staging.gc_ratio_global(c)[0] == np.float32(0.5)
    low = codes_of("acgtACGT")
    assert list(low[0]) == [4, 4, 4, 4, 0, 1, 2, 3]               # lowercase is N (onehot_dna.py:23-30)


def test_kat5_blur_single_hit_and_edges():
    X0 = np.zeros((1, 20), dtype=np.float32)
    X0[0, 10] = 5.0
    X = profile.blur(X0, 2)
    assert np.allclose(X[0, 8:13], 1.0) and X[0].sum() == pytest.approx(5.0)
    X0[0, :] = 0
    X0[0, 0] = 5.0
    X = profile.blur(X0, 2)
    assert X[0, 0] == 0 and X[0, 1] == 0 and X[0, 2] == pytest.approx(1.0)   # first/last `margin` positions are zero
    assert np.array_equal(profile.blur(X0, 0), X0.astype(np.float64))


def test_kat6_correlation():
    X = np.array([[0.], [0.], [1.], [1.]])
    Y = np.array([[1.], [2.], [3.], [4.]])
    r = profile.positional_correlation(X, Y)
    assert r[0, 0] == pytest.approx(0.894427, abs=1e-6)
    r = profile.positional_correlation(X, Y, np.full(4, 0.5))     # constant z -> NaN -> 0 after nan_to_num
    assert np.isnan(r[0, 0]) and np.nan_to_num(r.astype(np.float32), 0.0)[0, 0] == 0.0
    rng = np.random.default_rng(3)
    x, y, z = rng.normal(size=50), rng.normal(size=50), rng.normal(size=50)
    r = profile.positional_correlation(x[:, None], y[:, None], z)[0, 0]
    cc = np.corrcoef([x, y, z])
    want = (cc[0, 1] - cc[0, 2] * cc[1, 2]) / np.sqrt((1 - cc[0, 2] ** 2) * (1 - cc[1, 2] ** 2))
    assert r == pytest.approx(want, rel=1e-10)


def test_kat7_percentiles_and_pvalues():
    r = np.array([[0.15, -.2, -.1, .1, .2]], dtype=np.float32)
    cols = profile.positional_r_columns(r, confidence_interval_pct=95)
    assert cols['permutation_full_positional_r_lower'][0] == pytest.approx(-0.1925, abs=1e-6)
    assert cols['permutation_full_positional_r_upper'][0] == pytest.approx(0.1925, abs=1e-6)
    assert cols['permutation_full_positional_r_pval'][0] == 0.5
    assert cols['permutation_semi_positional_r_pval'][0] == 0.5
    assert list(cols.keys()) == profile.POSITIONS_COLUMNS[:15]


@pytest.mark.parametrize("grid", GRIDS)
def test_kat8_short_motif_threshold_above_max(grid):
    # a 5-mer's best score has probability 4^-5 > 1e-4 -> thr = max + one grid step -> no matches (quirk vi)
    pwm = consensus_pwm("ACGTA")
    W = moods.log_odds(pwm, BG, 1e-4)
    thr = moods.threshold_from_p(W, BG, 1e-4, grid)
    imax = int(sum(np.trunc(grid * W.max(axis=0) + 0.5)))
    assert thr == pytest.approx((imax + 1) / grid)
    # whether anything can still match depends on the integer rounding of the column maxima: 5 * 1.38622 = 6.9311;
    # grid 1000: 5 * 1386 + 1 = 6931 -> thr 6.931 < 6.9311, the exact consensus still passes, by < 1e-3;
    # grid 2000: 5 * 2772 + 1 = 13861 -> thr 6.9305, it passes by < 1e-3 as well
    X0 = profile.scan(codes_of("GG" + "ACGTA" + "GG"), W.astype(np.float32), None, thr)
    assert list(np.nonzero(X0[0])[0]) == [4] and 0 < X0[0, 4] < 1e-3
    # the shipped 5-mer 'Tgif1,Tgif2 Homeobox' has max 5.679 < thr -> no position can match
    from mepp_b200 import synth
    tg = synth.load_motif_library('homer_clustered')['Tgif1,Tgif2 Homeobox']
    Wt = moods.log_odds(tg, BG, 1e-4)
    assert Wt.shape[1] == 5 and moods.threshold_from_p(Wt, BG, 1e-4, grid) > Wt.max(axis=0).sum()


def test_kat9_dp_equals_bruteforce_on_random_motifs():
    rng = np.random.default_rng(7)
    for m in (4, 6, 8, 9):
        pwm = rng.dirichlet(np.full(4, 0.4), size=m).T
        W = moods.log_odds(pwm, BG, 1e-4)
        for p in (1e-4, 1e-3, 1e-2):
            assert moods.threshold_from_p(W, BG, p) == moods.threshold_bruteforce(W, BG, p)


@pytest.mark.parametrize("grid,table", [
    (1000.0, [6.393, 7.225, 7.127, 7.522, 6.318, 6.166, 7.482, 6.636, 6.329, 5.063]),
    (2000.0, [6.3905, 7.2255, 7.126, 7.5215, 6.319, 6.1655, 7.482, 6.6375, 6.3295, 5.062])])
def test_ohler_thresholds_frozen(grid, table):
    """Thresholds of the shipped ohler library under the restated MOODS spec, on both DP grids (the grid moves them by
    up to 2.5e-3: integer rounding of every matrix entry, summed over the motif)."""
    from mepp_b200 import synth
    lib = synth.load_motif_library('ohler')
    thr = [moods.threshold_from_p(moods.log_odds(v, BG, 1e-4), BG, 1e-4, grid) for v in lib.values()]
    assert thr == pytest.approx(table)


def test_n_base_uses_four_quarter_adds():
    rng = np.random.default_rng(11)
    W = rng.normal(size=(4, 6)).astype(np.float32)
    seq = "ACNGTA"
    X = profile._conv_valid_f32(codes_of(seq), W)
    acc = np.float32(0)
    for j, ch in enumerate(seq):
        if ch == 'N':
            for c in range(4):
                acc = np.float32(acc + np.float32(0.25) * W[c, j])
        else:
            acc = np.float32(acc + W["ACGT".index(ch), j])
    assert X[0, 0] == acc


def test_parsers_and_filter():
    txt = ">m1 NAME\n1 2 3\n1 1 1\n2 1 0\n0 0 0\n>m2\n1 0\n0 1\n0 0\n0 0\n>m1 NAME\n4 0 0\n0 4 0\n0 0 4\n0 0 0\n"
    d = staging.parse_motifs(txt)
    assert list(d.keys()) == ["m1 NAME", "m2"] and d["m1 NAME"][0, 0] == 1.0     # later duplicate wins
    fa = ">s1 2.5 extra\nACGT\nAC\n>s2 1.0\nACGTNN\n>s3 7\nACG\n>s4\nACGTAC\n"
    seqs, scores, descs = staging.parse_scored_fasta(fa)
    assert seqs["s1"] == "ACGTAC" and scores == {"s1": 2.5, "s2": 1.0, "s3": 7.0} and "s4" not in descs
    rows = staging.filter_and_order({k: seqs[k] for k in ("s1", "s2", "s3")}, scores, descs, 20)
    assert [r[0] for r in rows] == ["s1"]                                        # s3 short, s2 33 % degenerate


def test_c_restatement_equals_numpy_oracle():
    from oracle import cref
    from mepp_b200 import synth
    ascii_u8, _ = synth.synth_sequences(200, 90, seed=1, n_rate=0.02, lower_rate=0.02)
    codes = staging.ascii_to_codes(ascii_u8)
    for pwm in synth.synth_motifs(4, seed=2, m_lo=4, m_hi=30):
        for o in ('+', '-', '+/-'):
            Wf, Wr, thr = profile.motif_weights(pwm, o)
            a, b = profile.scan(codes, Wf, Wr, thr), cref.scan(codes, Wf, Wr, thr)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
            for mg in (0, 2):
                assert np.array_equal(profile.blur(a, mg, dtype=np.float32), cref.blur(a, mg))


def test_accuracy_report_gates_pass_on_oracle_output_and_fail_on_a_perturbed_one():
    """oracle/checks.py (the accuracy block of bench.py): the oracle's own dense results satisfy every gate, a
    dropped match / a shifted r do not."""
    from oracle import checks, profile, staging
    from mepp_b200 import synth
    cfg = synth.make_config("cfg1", n_seq=300, n_motifs=2, n_perm=40)
    codes = staging.ascii_to_codes(cfg["ascii"])
    Y = staging.permuted_scores(cfg["scores"], cfg["perm_idx"])
    z = staging.gc_ratio_global(codes)
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    hit_dtype = np.dtype([('task', np.int32), ('seq', np.int32), ('pos', np.int32), ('x0', np.float32)])
    hits, rs, pos = [], [], {}
    for k, (pwm, o) in enumerate(tasks):
        Wf, Wr, thr = profile.motif_weights(pwm, o)
        X0 = profile.scan(codes, Wf, Wr, thr)
        s, p = np.nonzero(X0)
        h = np.zeros(s.size, dtype=hit_dtype)
        h['task'], h['seq'], h['pos'], h['x0'] = k, s, p, X0[s, p]
        hits.append(h)
        X = profile.blur(X0, cfg["margin"], dtype=np.float32)
        r = np.nan_to_num(profile.positional_correlation(X.astype(np.float64), Y.astype(np.float64),
                                                         z.astype(np.float64)), nan=0.0).astype(np.float32)
        rs.append(r)
        cols = profile.positional_r_columns(r)
        for name, v in cols.items():
            pos.setdefault(name, []).append(np.asarray(v))
    hits = np.concatenate(hits)
    r = np.stack(rs)
    pos = {k: np.stack(v) for k, v in pos.items()}
    rep = checks.accuracy_report(codes, Y, z, tasks, list(range(len(tasks))), cfg["margin"], hits, r, pos, threads=2)
    assert rep["match_sets_equal"] and rep["match_scores_bit_equal"] and rep["r_within_1e5_gate"]
    assert rep["pvalue_flips"] == 0 and rep["ci_max_abs_err"] < 1e-6 and rep["matches_checked"] == hits.size > 0
    bad = checks.accuracy_report(codes, Y, z, tasks, list(range(len(tasks))), cfg["margin"], hits[1:], r * 1.001, pos,
                                 threads=2)
    assert not bad["match_sets_equal"] and not bad["r_within_1e5_gate"]

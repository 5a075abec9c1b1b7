"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side
threshold code agrees with the oracle, host statistics agree with numpy / the oracle, and the
staging layer keeps the reference's semantics.  No GPU compute is called here."""
import io
import os
import re

import numpy as np
import pytest

from oracle import moods, profile, staging

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mepp_b200.h")).read()
    declared = set(re.findall(r"\b(mepp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib.lib, name), f"{name} declared in include/mepp_b200.h but not exported"
    assert declared == set(lib.SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.lib.mepp_abi_version() == 1


def test_device_entry_points_fail_loudly_without_gpu(lib):
    import ctypes as C
    if lib.device_ok():
        pytest.skip("a GPU is present")
    rc = lib.lib.mepp_pack_dna(C.c_void_p(16), 4, 8, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), None)
    assert rc != 0 and "no sm_100 CUDA device" in lib.last_error()
    from mepp_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()
    from mepp_b200 import utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.force_cpu_only()


def test_host_threshold_matches_oracle_on_shipped_libraries(lib):
    from mepp_b200 import motif_layers, synth
    for libname in ("ohler", "homer_clustered"):
        for key, pwm in synth.load_motif_library(libname).items():
            for o in ('+', '-', '+/-'):
                tab = motif_layers.motif_matrix_to_scan_table(pwm, o)
                Wf, Wr, thr = profile.motif_weights(pwm, o)
                m = pwm.shape[1]
                assert tab['thr'] == thr, (key, o)
                assert tab['bias'] == np.float32(-thr)
                assert np.array_equal(tab['lut'][:m, :, 0], Wf.T)
                assert tab['n_filters'] == (2 if Wr is not None else 1)
                if Wr is not None:
                    assert np.array_equal(tab['lut'][:m, :, 1], Wr.T)
                assert (tab['lut'][m:] == 0).all()


def test_host_threshold_honours_flags_only_when_asked(lib):
    from mepp_b200 import motif_layers, synth
    pwm = list(synth.load_motif_library("ohler").values())[0]
    a = motif_layers.motif_matrix_to_scan_table(pwm, '+', pseudocount=0.01, pvalue=0.001)
    b = motif_layers.motif_matrix_to_scan_table(pwm, '+')
    assert a['thr'] == b['thr']                       # reference quirk: flags never reach MOODS (core.py:513-534)
    c = motif_layers.motif_matrix_to_scan_table(pwm, '+', pseudocount=0.01, pvalue=0.001, honour_flags=True)
    W = moods.log_odds(pwm, moods.flat_bg(), 0.01)
    assert c['thr'] == moods.threshold_from_p(W, moods.flat_bg(), 0.001) and c['thr'] < b['thr']
    with pytest.raises(ValueError):
        motif_layers.motif_matrix_to_scan_table(np.full((4, 40), 0.25), '+')


def test_log_odds_and_threshold_wrappers(lib):
    from mepp_b200 import motif_layers
    rng = np.random.default_rng(1)
    pwm = rng.dirichlet(np.ones(4), size=9).T
    W = motif_layers.log_odds(pwm)
    assert np.allclose(W, moods.log_odds(pwm, moods.flat_bg(), 1e-4), rtol=0, atol=1e-15)
    for p in (1e-4, 1e-3):
        assert motif_layers.threshold_from_p(W, pvalue=p) == moods.threshold_bruteforce(W, moods.flat_bg(), p)


def test_lerp_matches_numpy_percentile():
    from mepp_b200 import stats_host
    rng = np.random.default_rng(2)
    for P in (4, 100, 1000, 777):
        r = rng.normal(scale=0.01, size=(13, P)).astype(np.float32)
        for ci in (95, 90, 50):
            q = stats_host.percentile_q(ci)
            want = np.percentile(r, q, axis=-1)
            s = np.sort(r, axis=-1)
            for i, qq in enumerate(q):
                k, g = stats_host.virtual_index(P, qq)
                got = stats_host.lerp_f32(s[:, k], s[:, min(k + 1, P - 1)], g)
                assert np.array_equal(got, want[i]), (P, ci, i)


def test_finalize_positions_matches_oracle_columns():
    """Feed finalize_positions exact order statistics / counts computed on the host from a random r."""
    from mepp_b200 import stats_host
    rng = np.random.default_rng(3)
    T, L, P, N = 3, 41, 60, 500
    r = rng.normal(scale=0.05, size=(T, L, 1 + P)).astype(np.float32)
    r[0, 5, :] = 0.0
    count = rng.integers(0, 9, size=(T, L))
    q = stats_host.percentile_q(95)
    k_lo, _ = stats_host.virtual_index(P, q[0]); k_hi, _ = stats_host.virtual_index(P, q[1])
    ks_lo, _ = stats_host.virtual_index(L * P, q[0]); ks_hi, _ = stats_host.virtual_index(L * P, q[1])
    s = np.sort(r[:, :, 1:], axis=-1)
    full_os = np.stack([s[..., k_lo], s[..., k_lo + 1], s[..., k_hi], s[..., min(k_hi + 1, P - 1)]], axis=-1)
    full_cnt = (np.abs(r[:, :, 1:]) >= np.abs(r[:, :, 0:1])).sum(-1)
    flat = np.sort(r[:, :, 1:].reshape(T, -1), axis=-1)
    semi_os = np.stack([flat[:, ks_lo], flat[:, ks_lo + 1], flat[:, ks_hi], flat[:, min(ks_hi + 1, L * P - 1)]], axis=-1)
    semi_cnt = (np.abs(r[:, :, 1:]).reshape(T, 1, -1) >= np.abs(r[:, :, 0])[:, :, None]).sum(-1)
    a = np.abs(r).astype(np.float64)
    integral = (0.5 * (a[:, 1:] + a[:, :-1])).sum(axis=1)
    cols = stats_host.finalize_positions(r[:, :, 0], full_os, full_cnt, semi_os, semi_cnt, integral, count, N, P, 2)
    for t in range(T):
        want = profile.positional_r_columns(r[t])
        lo, up, pv = profile.parametric_ci(want['positional_r'], N)
        want.update(positional_r_lower=lo, positional_r_upper=up, positional_r_pval=pv,
                    count=count[t].astype(float), smoothed_count=profile.smooth_counts(count[t].astype(float), 5))
        assert list(cols.keys()) == profile.POSITIONS_COLUMNS
        for k in profile.POSITIONS_COLUMNS:
            a_, b_ = np.asarray(cols[k][t], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
            if 'integral' in k:
                assert np.allclose(a_, b_, rtol=1e-6, atol=1e-7), k     # float32 trapz summation order
            else:
                assert np.array_equal(a_, b_), k


def test_smooth_counts_matches_pandas_rolling():
    import pandas as pd
    from mepp_b200 import stats_host
    c = np.random.default_rng(4).integers(0, 50, size=37).astype(float)
    for w in (3, 5, 11):
        want = (pd.DataFrame(dict(position=np.arange(37), count=c)).set_index('position')
                .rolling(w, center=True).mean().reset_index().ffill().bfill())['count'].to_numpy()
        assert np.allclose(stats_host.smooth_counts(c, w), want, rtol=0, atol=1e-12)
        assert np.allclose(profile.smooth_counts(c, w), want, rtol=0, atol=1e-12)


def test_io_parsers_match_oracle_and_reference_rules():
    from mepp_b200 import io as mio
    txt = ">m1 NAME\n1 2 3\n1 1 1\n2 1 0\n0 0 0\n>m2\n1 0\n0 1\n0 0\n0 0\n>m1 NAME\n4 0 0\n0 4 0\n0 0 4\n0 0 0\n"
    d, cons = mio.motif_matrix_file_to_dicts(io.StringIO(txt))
    o = staging.parse_motifs(txt)
    assert list(d.keys()) == list(o.keys()) == ["m1 NAME", "m2"]
    for k in d:
        assert np.array_equal(d[k], o[k])
    assert cons["m1 NAME"] == "ACG" and cons["m2 m2"] == "AC"
    fa = ">s1 2.5 extra\nACGT\nAC\n>s2 1.0\nACGTNN\n>s3 7\nACG\n>s4\nACGTAC\n"
    assert mio.scored_fasta_file_to_dicts(io.StringIO(fa)) == staging.parse_scored_fasta(fa)


def test_filter_order_dataset_roundtrip(tmp_path):
    from mepp_b200 import io as mio, utils, permutations
    fa = "".join(f">s{i} {s}\n{seq}\n" for i, (s, seq) in enumerate(
        [(3.0, "ACGTAC"), (1.0, "ACGTNN"), (2.0, "acgtAC"), (0.5, "ACG"), (2.5, "TTTTTT")]))
    dicts = mio.scored_fasta_file_to_dicts(io.StringIO(fa))
    df = utils.order_scored_fasta_df(utils.filter_scored_fasta_df(utils.scored_fasta_dicts_to_df(*dicts), 50),
                                     random_state=0)
    assert list(df['sequence_id']) == ['s1', 's4', 's0']       # s3 short, s2 66 % degenerate (lowercase), sorted by score
    assert list(df.columns) == ['sequence_id', 'sequence', 'score', 'description', 'length', 'degenerate_pct', 'rank']
    ds = utils.append_gc_ratios_to_dataset(utils.scored_fasta_df_to_dataset(df))
    ds = permutations.add_permuted_scores_to_dataset(ds, num_permutations=7, seed=1)
    assert ds.num_permutations == 7 and ds.has_gc and ds.sequence_length == 6 and len(ds.element_spec) == 3
    assert sorted(ds.perm_idx[3]) == [0, 1, 2]
    assert permutations.add_permuted_scores_to_dataset(ds, num_permutations=0) is ds
    path = str(tmp_path / "dataset")
    mio.save_dataset(ds, path, delete_existing=True)
    assert os.path.exists(path + ".element_spec.p")
    ds2 = mio.load_dataset(path)
    assert np.array_equal(ds2.ascii, ds.ascii) and np.array_equal(ds2.perm_idx, ds.perm_idx) and ds2.has_gc


def test_vectorised_degenerate_filter_equals_the_per_sequence_rule():
    """filter_scored_fasta_df counts the characters outside upper-case ACGT of all sequences at once (one translate over
    the joined bytes); it has to give exactly what onehot_dna.get_degenerate_pct -- the reference's per-sequence rule,
    onehot_dna.py:41-53 -- gives, for lower case, N, IUPAC codes, digits and characters outside latin-1 alike."""
    import pandas as pd
    from mepp_b200 import utils, onehot_dna
    rng = np.random.default_rng(3)
    letters = list("ACGTacgtNnRYKM-*0 ") + ["\u00e9", "\u4e2d"]
    seqs = ["".join(rng.choice(letters, size=40, p=None)) for _ in range(200)]
    seqs[0] = "ACGT" * 10
    df = pd.DataFrame({'sequence_id': [f"s{i}" for i in range(len(seqs))], 'sequence': seqs,
                       'score': rng.normal(size=len(seqs)), 'description': [("s", "0")] * len(seqs)})
    out = utils.filter_scored_fasta_df(df, degenerate_pct_thresh=100)
    assert len(out) == len(seqs)
    want = np.array([onehot_dna.get_degenerate_pct(s) for s in out['sequence']])
    assert np.array_equal(out['degenerate_pct'].to_numpy(), want)
    assert want[0] == 0.0 and want.max() > 50.0
    for thr in (0, 25.0, 60.0):
        kept = utils.filter_scored_fasta_df(df, degenerate_pct_thresh=thr)
        assert list(kept['sequence']) == [s for s in seqs if onehot_dna.get_degenerate_pct(s) <= thr]


def test_synthetic_generator_is_deterministic_and_shaped():
    from mepp_b200 import synth
    a = synth.make_config("cfg1", n_motifs=2)
    b = synth.make_config("cfg1", n_motifs=2)
    assert np.array_equal(a["ascii"], b["ascii"]) and np.array_equal(a["perm_idx"], b["perm_idx"])
    assert a["ascii"].shape == (1000, 201) and a["perm_idx"].shape == (100, 1000) and len(a["tasks"]) == 4
    assert (np.diff(a["scores"]) >= 0).all()
    frac_n = (a["ascii"] == ord('N')).mean()
    assert 0.0005 < frac_n < 0.005
    assert len(synth.load_motif_library("homer_clustered")) == 162       # 164 headers, 2 duplicate keys


def test_layout_decoders_roundtrip():
    from mepp_b200 import layout
    rng = np.random.default_rng(0)
    m_rows, n_pad = 200, 64
    MT, KB = 2, 2
    data = rng.integers(0, 1000, size=(MT, KB, 2, 4, 128, 8)).astype(np.float16)
    flags = np.zeros((MT, 4), dtype=np.uint32)
    flags[0, 0] = 0b0110          # tile 0: k-block 0 half 1, k-block 1 half 0
    flags[1, 0] = 0b1000          # tile 1: k-block 1 half 1
    raw = np.concatenate([data.ravel().view(np.uint8), flags.ravel().view(np.uint8)])
    hi, lo, nz = layout.decode_x_planes(raw, m_rows, n_pad, with_flags=True)
    assert hi[130, 41] == np.float32(data[1, 1, 0, 1, 2, 1]) and lo[5, 7] == np.float32(data[0, 0, 1, 0, 5, 7])
    assert nz.tolist() == [[[False, True], [True, False]], [[False, False], [False, True]]]


def test_cli_flags_match_the_reference():
    """Every flag of the reference's `mepp` command (mepp/cli.py:12-456) is accepted."""
    from click.testing import CliRunner
    from mepp_b200 import cli
    flags = {o for p in cli.main.params for o in p.opts + p.secondary_opts}
    expected = {'--fa', '--motifs', '--out', '--center', '--dgt', '--perms', '--batch', '--jobs', '--keepdata',
                '--orientations', '--margin', '--pcount', '--pval', '--bg', '--ci', '--sigma', '--cmap', '--smoothing',
                '--width', '--height', '--formats', '--dpi', '--gjobs', '--nogpu', '--attempts', '--minwait',
                '--maxwait', '--cmethod', '--cmetric', '--tdpi', '--tformat', '--mtmethod', '--mtalpha',
                '--thoroughmt', '--non-thoroughmt'}
    assert expected <= flags
    # names, destinations, defaults, flag-ness and required-ness of every option equal the reference's
    # (tests/golden/reference_cli_options.json, made from mepp/cli.py by tests/golden/make_cli_fixture.py)
    import json
    import multiprocessing
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_cli_options.json')) as f:
        ref_opts = json.load(f)
    assert len(ref_opts) == 35
    ours = {o: p for p in cli.main.params for o in p.opts + p.secondary_opts}
    for ro in ref_opts:
        for flag in ro['flags']:
            p = ours[flag]
            assert p.name == ro['dest'], flag
            assert bool(p.required) == ro['required'], flag
            assert bool(getattr(p, 'is_flag', False)) == (ro['is_flag'] or ro['flag_value'] is not None), flag
            if ro['flag_value'] is not None:
                assert p.flag_value == eval(ro['flag_value']), flag
            if ro['default'] is not None:
                want = eval(ro['default'], {'multiprocessing': multiprocessing})
                got = p.default
                assert (tuple(got) if isinstance(want, tuple) else got) == want, (flag, got, want)
            elif not ro['is_flag']:
                assert p.default is None, flag
    res = CliRunner().invoke(cli.main, ['--help'])
    assert res.exit_code == 0 and '--help' in res.output
    res = CliRunner().invoke(cli.main, ['--fa', 'x', '--motifs', 'y', '--out', 'z', '--nogpu'])
    assert res.exit_code != 0 and 'no CPU fallback' in res.output


def test_batch_paths_and_slugs():
    from mepp_b200 import batch
    assert batch.orientation_to_filepath == {'+': 'fwd', '-': 'rev', '+/-': 'fwd-rev'}
    assert batch.motif_id_and_orientation_to_filepath('PAX5,Pax8 Paired,Homeobox', '+/-') == \
        'pax5-pax8-paired-homeobox/orientation_fwd-rev'
    assert batch.slugify('ohler_motif_2 DRE') == 'ohler-motif-2-dre'


def test_filter_table_is_conservative(lib):
    """The scan kernel's integer pre-filter (host.cu build_filter_table) must never reject a window whose
    exact float32 score passes the threshold: check the bound on random windows, N bases included."""
    from mepp_b200 import motif_layers, synth
    rng = np.random.default_rng(9)
    for pwm in synth.synth_motifs(6, seed=13, m_lo=3, m_hi=32) + list(synth.load_motif_library('ohler').values())[:3]:
        for o in ('+', '-', '+/-'):
            tab = motif_layers.motif_matrix_to_scan_table(pwm, o)
            m, nf = tab['m'], tab['n_filters']
            q = tab['qtab'].reshape(16, 64)
            qinit = int(tab['qthr'][0])
            syms = rng.integers(0, 5, size=(4000, m))
            syms[:2000][syms[:2000] == 4] = rng.integers(0, 4, size=(syms[:2000] == 4).sum())   # half without N
            # plant near-consensus windows so that real matches are present
            best = np.argmax(tab['lut'][:m, :, 0], axis=1)
            syms[:200] = best
            syms[100:200, rng.integers(0, m, 100)] = rng.integers(0, 4, 100)
            pad = np.zeros((syms.shape[0], 32), dtype=np.int64)
            pad[:, :m] = syms
            idx = pad[:, 0::2] | (pad[:, 1::2] << 3)
            packed = q[np.arange(16)[None, :], idx].astype(np.int64).sum(axis=1) + qinit
            passes = (packed & 0x80008000) != 0
            # exact canonical float32 scores
            W = tab['lut'][:m]
            hit = np.zeros(len(syms), dtype=bool)
            for f in range(nf):
                acc = np.zeros(len(syms), dtype=np.float32)
                for j in range(m):
                    sj = syms[:, j]
                    normal = acc + W[j, np.minimum(sj, 3), f]
                    an = acc
                    for c in range(4):
                        an = an + np.float32(0.25) * W[j, c, f]
                    acc = np.where(sj == 4, an, normal).astype(np.float32)
                hit |= (acc + tab['bias']) > 0
            assert not (hit & ~passes).any(), "pre-filter rejected a true match"
            assert hit[:100].all() or tab['thr'] > W[:, :, 0].max(axis=1).sum()      # consensus windows do match


def test_heatmap_from_match_list_equals_dense_reduction():
    """clip + block-max of the score matrix from its non-zero entries only."""
    from mepp_b200 import heatmap
    rng = np.random.default_rng(5)
    N, L = 517, 203
    X0 = np.zeros((N, L), dtype=np.float32)
    k = 400
    X0[rng.integers(0, N, k), rng.integers(0, L, k)] = rng.gamma(2.0, 1.0, k).astype(np.float32)
    seq, pos = np.nonzero(X0)
    for (pw, ph, sigma) in ((100, 60, 0.5), (1000, 1000, 1.0), (37, 11, None), (203, 517, 2.0)):
        a = heatmap.compress_hits_to_pixels(seq, pos, X0[seq, pos], X0.shape, pw, ph, sigma)
        b = heatmap.compress_dense_to_pixels(X0, pw, ph, sigma)
        assert a.shape == b.shape and np.array_equal(a, b), (pw, ph, sigma)
    assert heatmap.get_aspect_tensor((100000, 1000), 800, 600) == (167, 1)
    empty = heatmap.compress_hits_to_pixels([], [], [], (10, 10), 5, 5)
    assert empty.shape == (5, 5) and not empty.any()


def test_multipletests_known_answers():
    from mepp_b200.summary import multipletests
    p = np.array([0.01, 0.04, 0.03, 0.005, 0.8])
    rej, adj = multipletests(p, alpha=0.05, method='fdr_bh')
    assert np.allclose(adj, [0.025, 0.05, 0.05, 0.025, 0.8]) and rej.tolist() == [True, True, True, True, False]
    _, adj = multipletests(p, alpha=0.05, method='bonferroni')
    assert np.allclose(adj, [0.05, 0.2, 0.15, 0.025, 1.0])
    _, adj = multipletests(p, alpha=0.05, method='holm')
    assert np.allclose(adj, [0.04, 0.09, 0.09, 0.025, 0.8])
    _, adj = multipletests(p, alpha=0.05, method='fdr_by')
    assert np.allclose(adj, np.minimum(np.array([0.025, 0.05, 0.05, 0.025, 0.8]) * (1 + 1/2 + 1/3 + 1/4 + 1/5), 1.0))
    rng = np.random.default_rng(0)
    q = np.concatenate([rng.uniform(0, 1e-4, 30), rng.uniform(0, 1, 300)])
    for method in ('fdr_tsbky', 'fdr_tsbh', 'sidak', 'holm-sidak', 'simes-hochberg'):
        rej, adj = multipletests(q, alpha=0.01, method=method)
        assert adj.shape == q.shape and (adj <= 1).all()
        assert method.startswith('fdr_ts') or (adj >= q - 1e-15).all()   # two-stage scales by m0/m < 1
        assert (np.diff(adj[np.argsort(q)]) >= -1e-12).all()           # monotone in p
    # two-stage BKY is never more conservative than BH at level alpha / (1 + alpha) when something is rejected
    _, bky = multipletests(q, alpha=0.01, method='fdr_tsbky')
    _, bh = multipletests(q, alpha=0.01, method='fdr_bh')
    assert (bky <= bh * 1.01 + 1e-12).all()


def test_minmax_stats_table():
    import pandas as pd
    from mepp_b200 import stats_host, summary
    rng = np.random.default_rng(1)
    T, L, P = 4, 30, 20
    r0 = rng.normal(scale=0.05, size=(T, L)).astype(np.float32)
    store = stats_host.alloc_positions(T, L, P)
    full_os = np.sort(rng.normal(scale=0.05, size=(T, L, 4)).astype(np.float32), axis=-1)
    stats_host.finalize_into(store, 0, r0, full_os, rng.integers(0, P, (T, L)), np.sort(rng.normal(size=(T, 4)).astype(np.float32)),
                             rng.integers(0, L * P, (T, L)), np.abs(rng.normal(size=(T, 1 + P))), rng.integers(0, 3, (T, L)), 100, 2)
    cols = stats_host.positions_view(store)
    dfs = {f"m{t}": pd.DataFrame({k: np.array(v[t]) for k, v in cols.items()}) for t in range(T)}
    st = summary.get_minmax_stats(dfs["m0"])
    i = int(np.argmax(r0[0]))
    assert st['max_r'] == r0[0, i] and st['max_r_pos'] == i - L // 2
    assert st['max_r_pval'] == dfs["m0"]['permutation_positional_r_pval'][i]
    for thorough in (True, False):
        tab = summary.get_minmax_stats_df(dfs, mt_alpha=0.05, thorough_mt=thorough)
        assert list(tab['motif_id']) == list(dfs) and {'extreme_r_padj', 'extreme_r_sig', 'integral_r_padj'} <= set(tab.columns)
        assert (tab['extreme_r_padj'] >= tab['extreme_r_pval'] - 1e-15).all()


def test_prepare_tables_matches_per_task_tables():
    """mepp_prepare_tables (everything but the threshold, whole task list in one call) against mepp_motif_table per
    task: float32 weights, filter counts and quantised pair tables identical; unit / twin maps as documented."""
    from mepp_b200 import synth, motif_layers
    cfg = synth.make_config("cfg1")
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]] + [(cfg["tasks"][0][1], '-'), (cfg["tasks"][2][1], '+')]
    for kw in (dict(), dict(honour_flags=True, pseudocount=0.01, pvalue=1e-3, bg=[0.3, 0.2, 0.2, 0.3])):
        prep = motif_layers.prepare_tables(tasks, **kw)
        ref = motif_layers.scan_tables(tasks, **kw)
        for t, tb in enumerate(ref):
            assert np.array_equal(prep['lut'][t], tb['lut']) and prep['nfilt'][t] == tb['n_filters']
            assert np.array_equal(prep['qtab'][t], tb['qtab']) and prep['mlen'][t] == tb['m']
        # '+' and '+/-' of a motif share a unit; '-' has its own; the first '+' of a motif is the twin of its '+/-'
        u = prep['unit_of_task']
        assert u[0] == u[1] and u[-2] != u[0] and u[-1] == u[2]
        assert prep['twin_of'][1] == 0 and prep['twin_of'][0] == -1 and prep['twin_of'][3] == 2
        # the DP inputs reproduce the host threshold when run through the host DP semantics
        R = prep['dp_par'][:, 7]
        assert (R >= 0).all() and (prep['smat'] >= 0).all()


def test_match_driven_blur_is_bit_identical_to_the_dense_blur():
    """The lean scan builds the blurred entries from the sorted match list (bucket_entries_kernel in csrc/scan.cu);
    tools/proto_match_blur.py is its numpy form: identical bits to the dense float32 blur of core.py:59-145
    (sequential sum in position order, IEEE division), for isolated and overlapping matches, zeroed edges included,
    and against the oracle's blur for margins 0 ... 16."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("proto_match_blur", os.path.join(here, "..", "tools", "proto_match_blur.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    from oracle import profile
    for margin in (0, 1, 2, 5, 16):
        assert proto._check(margin=margin, seed=margin) > 0
        rng = np.random.default_rng(100 + margin)
        X0 = np.where(rng.random((40, 97)) < 0.01, rng.gamma(2.0, 1.0, size=(40, 97)), 0.0).astype(np.float32)
        dense = proto.dense_blur_f32(X0, margin)
        assert np.allclose(dense, profile.blur(X0, margin, dtype=np.float64), rtol=3e-7, atol=0)


def test_tc_scan_pipeline_protocol_has_no_deadlock():
    """Barrier protocol of the tensor-core scan kernel (csrc/tcscan.cu), simulated on the CPU with random scheduling and
    randomly delayed asynchronous arrivals (tools/sim_tc_pipeline.py): every mix of N-tile widths runs to completion."""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location(
        "sim_tc_pipeline", os.path.join(os.path.dirname(__file__), "..", "tools", "sim_tc_pipeline.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    rnd = random.Random(11)
    cases = [[256, 96], [64], [256], [192, 160], [256, 256, 128], [96, 64, 32]]
    cases += [[32 * rnd.randint(1, 8) for _ in range(rnd.randint(1, 8))] for _ in range(4)]
    for bns in cases:
        for seed in range(2):
            assert sim.run(40, bns, seed=seed), (bns, seed)
    # tc_scan4i_kernel (one issuer per accumulator stage; an issuer releases only the stream tiles it has an N-tile of, so
    # the ring's "empty" barriers count min(n_tiles, 4) arrivals), whole stages and the half-stage form
    for n_tiles in range(1, 17):
        for split in (1, 2):
            assert sim.run4i(40, n_tiles, split=split, seed=n_tiles), (n_tiles, split)
    # the simulation has teeth: with four expected releases per stream tile three N-tiles per tile starve the producer
    assert not sim.run4i(40, 3, n_release=4, max_steps=200_000)


def test_native_tsv_writer_is_byte_identical_to_pandas(tmp_path):
    """mepp_write_tsv (csrc/writer.cu) against DataFrame.to_csv(sep='\\t', index=False) -- how single.py:223-241
    writes positions_df.tsv / ranks_df.tsv: shortest digits per column precision, numpy's positional / scientific
    rule, NaN as the empty string, broadcast (stride 0) columns."""
    import pandas as pd
    from mepp_b200 import writer
    rng = np.random.Generator(np.random.PCG64(5))
    n = 4000
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-4, 9.999999e-5, 1e-5, 1.5e-7, 1e15, 9.999999999999998e15, 1e16, 1.2345e22,
                        np.nan, np.inf, -np.inf, 5e-324, 1.7976931348623157e308, 0.1, 1 / 3, 123456.789, 2.5e-300, 3.0])
    f64 = np.concatenate([special, rng.normal(size=n - special.size) * 10.0 ** rng.integers(-12, 20, n - special.size)])
    f32 = np.concatenate([special.astype(np.float32), (rng.normal(size=n - special.size) *
                                                       10.0 ** rng.integers(-10, 12, n - special.size)).astype(np.float32)])
    cols = {
        'positional_r': f32, 'permutation_full_positional_r_lower': f64,
        'position': np.arange(n) - n // 2, 'count': rng.integers(0, 50, n).astype(np.float64),
        'i32': rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype(np.int32),
        'integral_r': np.broadcast_to(np.float32(0.123456789), (n,)),
        'bcast64': np.broadcast_to(np.float64(-1.25e-9), (n,)),
        'u16': rng.integers(0, 65535, n).astype(np.uint16),
    }
    a, b = str(tmp_path / 'a.tsv'), str(tmp_path / 'b.tsv')
    writer.write_tsv(a, cols)
    pd.DataFrame({k: np.array(v) for k, v in cols.items()}).to_csv(b, sep='\t', index=False)
    assert open(a, 'rb').read() == open(b, 'rb').read()
    # density: uint16 on the wire, float64 in the reference's ranks_df (core.py:456-470)
    dens = rng.integers(0, 7, n).astype(np.uint16)
    scores = rng.normal(5, 2, n).astype(np.float32)
    writer.write_tsv(a, dict(rank=np.arange(n), score=scores, density=dens), float_uint16=('density',))
    pd.DataFrame(dict(rank=np.arange(n), score=scores, density=dens.astype(np.float64))).to_csv(b, sep='\t', index=False)
    assert open(a, 'rb').read() == open(b, 'rb').read()
    # empty table, and a dtype the native writer does not take (falls to pandas, same bytes by construction)
    writer.write_tsv(a, dict(x=np.zeros(0), y=np.zeros(0, dtype=np.int64)))
    assert open(a).read() == 'x\ty\n'
    writer.write_tsv(a, dict(s=np.array(['p', 'q'], dtype=object), v=np.array([1.5, 2.5])))
    assert open(a).read() == 's\tv\np\t1.5\nq\t2.5\n'


def test_libyaml_dump_equals_the_python_emitter(tmp_path):
    """writer.yaml_dump (libyaml) writes the text yaml.dump(..., default_flow_style=False, allow_unicode=True) does
    (single.py:280-302) for the parameter / filepath dictionaries of a task."""
    import yaml
    from mepp_b200 import writer
    from mepp_b200.single import _task_filepaths
    params = dict(dataset_filepath='out/dataset', motif_matrix_filepath='out/x/orientation_fwd/motif_matrix.npy',
                  motif_id='MA0001.1 AGL3 é', output_filepath='out/x', center=None, motif_orientation='+/-',
                  motif_margin=2, motif_pseudocount=0.0001, motif_pvalue=1e-4, bg=None, confidence_interval_pct=95,
                  motif_score_sigma=0.5, motif_score_cmap='gray', rank_smoothing_factor=5, figure_width=10,
                  figure_height=10, save_datasets=False, save_profile_data=True, save_motif_score_matrix=False,
                  save_figures=True, figure_formats=['png', 'svg'], figure_dpi=300, no_gpu=False,
                  bg_list=[0.25, 0.25, 0.25, 0.25])
    fps = _task_filepaths('out/x', 'out/dataset', 'out/x/motif_matrix.npy', ['png', 'svg'])
    for obj in (params, fps):
        p = str(tmp_path / 'o.yaml')
        writer.yaml_dump(obj, p)
        assert open(p, encoding='utf8').read() == yaml.dump(obj, Dumper=yaml.Dumper, default_flow_style=False,
                                                            allow_unicode=True)
        assert yaml.safe_load(open(p, encoding='utf8')) == obj


def test_task_writer_pool_writes_the_reference_layout_and_isolates_failures(tmp_path):
    """writer.TaskWriter: spawned writer processes produce the files of single.py:198-302 / 342-391 per task; a job
    that fails (an error recorded by the scheduler, or an unwritable path) yields retval 1 for that task only."""
    import pandas as pd
    from mepp_b200 import writer
    from mepp_b200.single import _task_filepaths
    rng = np.random.Generator(np.random.PCG64(9))
    N, L = 500, 60
    scores = np.sort(rng.normal(5, 2, N)).astype(np.float32)
    tw = writer.TaskWriter(scores, n_workers=2)
    jobs = []
    for i in range(3):
        outdir = str(tmp_path / f'm{i}' / 'orientation_fwd')
        fps = _task_filepaths(outdir, str(tmp_path / 'dataset'), f'{outdir}/motif_matrix.npy', ['png'])
        fps.pop('mepp_plot')
        pos = dict(positional_r=rng.normal(size=L).astype(np.float32), position=np.arange(L) - L // 2,
                   integral_r=np.broadcast_to(np.float32(0.5), (L,)), count=rng.integers(0, 9, L).astype(np.float64))
        job = dict(outdir=outdir, motif_matrix=np.full((4, 6), 0.25), wrapped_params=dict(motif_id=f'm{i}', center=None),
                   params=dict(motif_id=f'm{i}'), filepaths=fps, positions=pos,
                   density=rng.integers(0, 4, N).astype(np.uint16), motif_score_matrix=None, save_profile_data=True,
                   save_motif_score_matrix=False, error='motif longer than 32 taps' if i == 1 else None)
        jobs.append(job)
        tw.submit(job)
    res = tw.results()
    tw.close()
    assert [r[0] for r in res] == [0, 1, 0]
    assert 'motif longer than 32 taps' in open(res[1][2]).read() and open(res[0][2]).read() == ''
    for i in (0, 2):
        d = jobs[i]['outdir']
        for f in ('motif_matrix.npy', 'wrapped_params.yaml', 'wrapped_params.log', 'params.yaml', 'filepaths.yaml',
                  'positions_df.tsv', 'positions_df.pkl', 'ranks_df.tsv', 'ranks_df.pkl'):
            assert os.path.exists(os.path.join(d, f)), f
        ranks = pd.read_pickle(os.path.join(d, 'ranks_df.pkl'))
        assert list(ranks.columns) == ['rank', 'score', 'density'] and ranks['density'].dtype == np.float64
        tsv = pd.read_csv(os.path.join(d, 'ranks_df.tsv'), sep='\t')
        assert np.array_equal(tsv['density'].to_numpy(), jobs[i]['density'].astype(np.float64))
        assert np.array_equal(tsv['score'].to_numpy().astype(np.float32), scores)
        pos = pd.read_pickle(os.path.join(d, 'positions_df.pkl'))
        assert list(pos.columns) == list(jobs[i]['positions'])
        assert open(os.path.join(d, 'positions_df.tsv')).read() == pos.to_csv(sep='\t', index=False)
    assert not os.path.exists(os.path.join(jobs[1]['outdir'], 'positions_df.tsv'))


def test_permutation_source_row_ranges_equal_the_full_matrix(tmp_path):
    """permutations.PermutationSource: any row range equals the same rows of the whole (P, N) matrix whatever the
    number of host threads (chunked random streams), every row is a permutation, and a dataset that holds a source
    survives save_dataset / load_dataset as its seed."""
    from mepp_b200 import permutations, io as mio
    from mepp_b200.dataset import ScoredDataset
    N, P = 1000, 300
    full = permutations.make_permutation_indices(N, P, seed=11, threads=1)
    assert full.shape == (P, N) and full.dtype == np.int32
    assert (np.sort(full, axis=1) == np.arange(N)).all()
    src = permutations.PermutationSource(N, P, seed=11)
    assert np.array_equal(src.rows(0, P, threads=4), full)
    assert np.array_equal(src.rows(128, 260, threads=3), full[128:260])
    assert np.array_equal(src.rows(256, 999), full[256:])
    with pytest.raises(ValueError):
        src.rows(10, 20)
    assert not np.array_equal(permutations.make_permutation_indices(N, P, seed=12), full)
    ds = ScoredDataset(np.full((N, 8), ord('A'), dtype=np.uint8), np.arange(N, dtype=np.float32), perm_idx=src)
    assert ds.num_permutations == P
    mio.save_dataset(ds, str(tmp_path / 'ds'))
    back = mio.load_dataset(str(tmp_path / 'ds'))
    assert hasattr(back.perm_idx, 'rows') and np.array_equal(back.perm_idx.rows(64, 128), full[64:128])

"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bit-exact for match positions / scores / counts; profiles,
confidence intervals within 1e-5 relative (with a floor tied to the row scale); p-values
reported with their discrete flips."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import profile, staging  # noqa: E402


def _engine():
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    from mepp_b200.engine import Engine
    return Engine()


def _r_close(r_gpu, r_orc, rel=1e-5, floor_frac=1e-5):
    """|dr| <= rel*|r| + floor_frac * (row scale), row scale = max_c |r[l, c]|."""
    d = np.abs(r_gpu.astype(np.float64) - r_orc.astype(np.float64))
    scale = np.abs(r_orc).max(axis=-1, keepdims=True)
    return d <= rel * np.abs(r_orc) + floor_frac * scale + 1e-12


def _oracle_inputs(cfg):
    codes = staging.ascii_to_codes(cfg["ascii"])
    Y = staging.permuted_scores(cfg["scores"], cfg["perm_idx"]) if cfg["perm_idx"] is not None and len(cfg["perm_idx"]) \
        else cfg["scores"][:, None]
    return codes, Y, staging.gc_ratio_global(codes)


@pytest.fixture(scope="module")
def cfg1():
    from mepp_b200 import synth
    return synth.make_config("cfg1")


def test_pack_matches_oracle(cfg1):
    eng = _engine().stage(cfg1["ascii"], cfg1["scores"], cfg1["perm_idx"])
    N, L = cfg1["N"], cfg1["L"]
    codes = staging.ascii_to_codes(cfg1["ascii"])
    from mepp_b200 import layout
    dec = layout.decode_symbols(eng.sym_T.cpu().numpy(), N, L)
    assert np.array_equal(dec, codes)
    assert np.array_equal(eng.gc[:N].cpu().numpy(), staging.gc_ratio_global(codes))   # bit-exact float32
    assert (eng.gc[N:].cpu().numpy() == 0).all()


def test_config1_full_parity(cfg1):
    """BASELINE.json configs[0]: 1000 x 201 bp, ohler motifs, '+' and '+/-', 100 permutations, margin 2."""
    eng = _engine().stage(cfg1["ascii"], cfg1["scores"], cfg1["perm_idx"])
    tasks = [(pwm, o) for _, pwm, o in cfg1["tasks"]]
    batch = eng.profile(tasks, margin=cfg1["margin"], return_hits=True, return_r=True)
    codes, Y, z = _oracle_inputs(cfg1)
    flips = 0
    for t, (pwm, o) in enumerate(tasks):
        orc = profile.profile_task(codes, Y, z, pwm, o, margin=cfg1["margin"])
        X0 = batch.motif_score_matrix(t, cfg1["L"])
        assert np.array_equal(X0.view(np.uint32), orc["X0"].view(np.uint32)), f"task {t}: match scores not bit-exact"
        assert batch.thr[t] == orc["thr"]
        assert np.array_equal(batch.density[t], orc["ranks"]["density"])
        assert _r_close(batch.r[t], orc["r"]).all(), f"task {t}: r outside tolerance"
        pos = batch.positions
        assert list(pos.keys()) == profile.POSITIONS_COLUMNS
        for col in profile.POSITIONS_COLUMNS:
            a = np.asarray(pos[col][t], dtype=np.float64)
            b = np.asarray(orc["positions"][col], dtype=np.float64)
            if col.endswith("_pval") and col.startswith("permutation"):
                flips += int((np.abs(a - b) > 1e-12).sum())
                assert np.abs(a - b).max() <= 0.05
            elif col in ("position", "count", "smoothed_count"):
                assert np.array_equal(a, b), col
            elif col == "positional_r_pval":
                assert np.allclose(a, b, rtol=1e-3, atol=1e-9), col
            else:
                scale = max(np.abs(b).max(), 1e-30)
                assert np.abs(a - b).max() <= 1e-5 * scale + 2e-7, (col, np.abs(a - b).max(), scale)
    # p-values are discrete (k/P): a 1-ulp difference in r can flip a '>='; report, allow a handful
    print("permutation p-value flips over all tasks:", flips)
    assert flips <= 20


@pytest.mark.parametrize("orientation", ['+', '-', '+/-'])
@pytest.mark.parametrize("margin", [0, 1, 5])
def test_orientations_and_margins(orientation, margin):
    from mepp_b200 import synth
    rng = np.random.default_rng(5)
    ascii_u8, scores = synth.synth_sequences(333, 77, seed=99, n_rate=0.02, lower_rate=0.02)   # ragged N (not % 32)
    perm = synth.synth_permutations(333, 17, seed=99)
    motifs = synth.synth_motifs(5, seed=3, m_lo=4, m_hi=32)
    eng = _engine().stage(ascii_u8, scores, perm)
    tasks = [(m, orientation) for m in motifs]
    batch = eng.profile(tasks, margin=margin, return_hits=True, return_r=True)
    codes = staging.ascii_to_codes(ascii_u8)
    Y = staging.permuted_scores(scores, perm)
    z = staging.gc_ratio_global(codes)
    for t, (pwm, o) in enumerate(tasks):
        orc = profile.profile_task(codes, Y, z, pwm, o, margin=margin)
        X0 = batch.motif_score_matrix(t, 77)
        assert np.array_equal(X0.view(np.uint32), orc["X0"].view(np.uint32))
        assert _r_close(batch.r[t], orc["r"]).all()
        assert np.array_equal(np.asarray(batch.positions["count"][t]), orc["positions"]["count"])
        assert np.array_equal(np.asarray(batch.positions["smoothed_count"][t]), orc["positions"]["smoothed_count"])


def test_edge_cases_all_n_no_hits_and_no_permutations():
    from mepp_b200 import synth
    L, N = 64, 40
    ascii_u8 = np.full((N, L), ord('N'), dtype=np.uint8)
    ascii_u8[::2] = np.frombuffer(b"ACGT" * 16, dtype=np.uint8)
    scores = np.linspace(-1, 1, N).astype(np.float32)
    cons = np.zeros((4, 8)); cons[[0, 1, 2, 3, 0, 1, 2, 3], np.arange(8)] = 1.0
    uniform = np.full((4, 6), 0.25)
    eng = _engine().stage(ascii_u8, scores, None)            # P = 0: no permutation columns
    batch = eng.profile([(cons, '+'), (uniform, '+/-')], margin=2, return_hits=True, return_r=True)
    assert list(batch.positions.keys()) == ['positional_r', 'position', 'positional_r_lower', 'positional_r_upper',
                                            'positional_r_pval', 'count', 'smoothed_count']
    codes = staging.ascii_to_codes(ascii_u8)
    z = staging.gc_ratio_global(codes)
    for t, (pwm, o) in enumerate([(cons, '+'), (uniform, '+/-')]):
        orc = profile.profile_task(codes, scores[:, None], z, pwm, o, margin=2)
        assert np.array_equal(batch.motif_score_matrix(t, L).view(np.uint32), orc["X0"].view(np.uint32))
        assert _r_close(batch.r[t], orc["r"]).all()
    assert (batch.density[1] == 0).all() and (np.asarray(batch.positions['positional_r'][1]) == 0).all()   # NaN -> 0
    # all-N sequences never match; ACGT-repeat sequences match the consensus at every 4th position
    assert (batch.density[0][1::2] == 0).all() and (batch.density[0][::2] > 0).all()


def test_gemm_matches_float64_check_kernel():
    """tcgen05 GEMM vs the CUDA-core float64 contraction of the same fp16 planes, K = 4096 sequences."""
    from mepp_b200 import _lib, synth
    from mepp_b200.engine import _ptr
    ascii_u8, scores = synth.synth_sequences(4096 + 5, 130, seed=4, motifs=synth.synth_motifs(2, seed=8))
    perm = synth.synth_permutations(4101, 300, seed=4)
    eng = _engine().stage(ascii_u8, scores, perm)
    motifs = synth.synth_motifs(3, seed=8)
    eng.fuse_correlation = False
    eng.profile([(m, '+/-') for m in motifs], margin=2)
    m_rows = 3 * 130
    Sd = torch.empty((m_rows, eng.C), dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib.mepp_profile_gemm_check(_ptr(eng._bufs['x_planes']), _ptr(eng.y_planes), _ptr(Sd), m_rows,
                                                eng.C, eng.n_pad, None, m_rows, None))
    torch.cuda.synchronize()
    S = eng._bufs['S'][:m_rows * eng.ldS].view(m_rows, eng.ldS)[:, :eng.C].double()
    rowmax = Sd.abs().amax(dim=1, keepdim=True)
    ok = (S - Sd).abs() <= 1e-5 * rowmax + 1e-3
    assert bool(ok.all()), f"max err/rowmax {(((S - Sd).abs()) / rowmax.clamp_min(1e-30)).max().item():.3e}"


def test_reference_entry_point_returns_the_reference_tuple(cfg1):
    """dataset_and_motif_matrix_to_profile_data keeps core.py:490-632's signature and 5-tuple."""
    from mepp_b200 import core, utils, permutations
    from mepp_b200.dataset import ScoredDataset
    ds = ScoredDataset(cfg1["ascii"], cfg1["scores"])
    ds = permutations.add_permuted_scores_to_dataset(utils.append_gc_ratios_to_dataset(ds), num_permutations=100,
                                                    perm_indices=cfg1["perm_idx"])
    pwm = cfg1["tasks"][0][1]
    msm, positions_df, ranks_df, processed, smoothed = core.dataset_and_motif_matrix_to_profile_data(
        ds, pwm, orientation='+/-', motif_margin=2)
    codes, Y, z = _oracle_inputs(cfg1)
    orc = profile.profile_task(codes, Y, z, pwm, '+/-', margin=2)
    assert msm.shape == (cfg1["N"], cfg1["L"]) and np.array_equal(msm.view(np.uint32), orc["X0"].view(np.uint32))
    assert list(positions_df.columns) == profile.POSITIONS_COLUMNS
    assert list(ranks_df.columns) == ['rank', 'score', 'density']
    assert np.array_equal(ranks_df['density'].to_numpy(), orc["ranks"]["density"])
    assert np.array_equal(ranks_df['score'].to_numpy(), cfg1["scores"])
    sm = smoothed.motif_scores()
    assert np.allclose(sm, profile.blur(orc["X0"], 2, dtype=np.float32), atol=1e-6)


def test_run_mepp_writes_the_reference_layout(tmp_path):
    """mepp.run_mepp end to end on a small scored FASTA: directory layout of SURVEY.md 8b and numbers vs the oracle."""
    import pandas as pd
    import yaml
    from mepp_b200 import synth, mepp as mepp_mod, batch as batch_mod, permutations
    ascii_u8, scores = synth.synth_sequences(300, 120, seed=21)
    rng = np.random.default_rng(0)
    scores = np.round(rng.normal(5, 2, 300), 4).astype(np.float32)        # unsorted: run_mepp must order them
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i in range(300):
            f.write(f">seq{i:04d} {scores[i]:.4f}\n{ascii_u8[i].tobytes().decode()}\n")
        f.write(">short 1.0\nACGT\n")                                    # dropped: not the max length
    motifs = synth.synth_motifs(2, seed=5, m_lo=6, m_hi=10)
    mf = tmp_path / "motifs.txt"
    with open(mf, "w") as f:
        for k, pwm in enumerate(motifs):
            f.write(f">M{k} name{k}\n")
            for row in pwm:
                f.write("\t".join(f"{v:.6f}" for v in row) + "\n")
    out = tmp_path / "out"
    filepaths_df, html, cmap = mepp_mod.run_mepp(str(fa), str(mf), str(out), num_permutations=30, motif_margin=2,
                                                motif_orientations=['+', '+/-'], permutation_seed=7, shuffle_seed=1)
    assert list(filepaths_df.columns) == ['motif_id', 'orientation', 'outdir', 'params', 'log', 'retval', 'attempts']
    assert len(filepaths_df) == 4 and (filepaths_df['retval'] == 0).all()
    assert (out / "filepaths.tsv").exists() and (out / "filtered_data_df.tsv").exists()
    for o in ('fwd', 'fwd-rev'):                                           # mepp.py:285-297 (tabular part)
        tab = pd.read_csv(out / f"results_table_orientation_{o}.tsv", sep="\t")
        assert {'motif_id', 'extreme_r', 'extreme_r_padj', 'extreme_r_sig', 'max_r_pos',
                'integral_r_padj'} <= set(tab.columns) and len(tab) == 2
    assert (out / "dataset.element_spec.p").exists() and not (out / "dataset").exists()     # batch.py:187-188
    fdf = pd.read_csv(out / "filtered_data_df.tsv", sep="\t")
    assert len(fdf) == 300 and (np.diff(fdf['score'].to_numpy()) >= 0).all()
    d = out / "m0-name0" / "orientation_fwd-rev"
    for name in ("motif_matrix.npy", "wrapped_params.yaml", "wrapped_params.log", "params.yaml", "filepaths.yaml",
                 "positions_df.tsv", "positions_df.pkl", "ranks_df.tsv", "ranks_df.pkl"):
        assert (d / name).exists(), name
    assert yaml.safe_load(open(d / "params.yaml"))['motif_orientation'] == '+/-'
    # numbers: rebuild the oracle inputs exactly as run_mepp staged them
    seq_sorted = np.stack([np.frombuffer(s.encode(), dtype=np.uint8) for s in fdf['sequence']])
    codes = staging.ascii_to_codes(seq_sorted)
    sc = fdf['score'].to_numpy().astype(np.float32)
    perm = permutations.make_permutation_indices(300, 30, seed=7)
    Y = staging.permuted_scores(sc, perm)
    z = staging.gc_ratio_global(codes)
    # file parsing divides counts by the column sum, so use the parsed matrix saved next to the results
    pwm = np.load(d / "motif_matrix.npy")
    orc = profile.profile_task(codes, Y, z, pwm, '+/-', margin=2)
    pos = pd.read_pickle(d / "positions_df.pkl")
    assert list(pos.columns) == profile.POSITIONS_COLUMNS
    assert np.array_equal(pos['count'].to_numpy(), orc['positions']['count'])
    assert _r_close(pos['positional_r'].to_numpy()[:, None], orc['r'][:, :1]).all()
    ranks = pd.read_pickle(d / "ranks_df.pkl")
    assert np.array_equal(ranks['density'].to_numpy(), orc['ranks']['density'])


def test_run_batch_isolates_a_poisoned_motif(tmp_path):
    """One motif the kernels cannot take (40 positions > 32) and one non-finite matrix fail ALONE -- retval 1 and a
    reason in wrapped_params.log for their rows of filepaths_df, like a failed child process of the reference
    (single.py:392-406, batch.py:174-185); every other task runs and its tables are what pandas would write."""
    import pandas as pd
    from mepp_b200 import synth, batch as batch_mod, utils as mutils, permutations
    from mepp_b200.dataset import ScoredDataset
    ascii_u8, scores = synth.synth_sequences(400, 150, seed=33)
    ds = permutations.add_permuted_scores_to_dataset(
        mutils.append_gc_ratios_to_dataset(ScoredDataset(ascii_u8, scores)), num_permutations=25, seed=3)
    good = synth.synth_motifs(3, seed=8, m_lo=6, m_hi=12)
    long_pwm = np.full((4, 40), 0.25)
    nan_pwm = good[0].copy()
    nan_pwm[0, 0] = np.nan
    motifs = {'G0 a': good[0], 'TOO_LONG x': long_pwm, 'G1 b': good[1], 'BAD nan': nan_pwm, 'G2 c': good[2]}
    fdf = batch_mod.run_batch(ds, motifs, str(tmp_path / 'out'), motif_orientations=['+', '+/-'], motif_margin=2)
    assert len(fdf) == 10
    bad = fdf['motif_id'].isin(['TOO_LONG x', 'BAD nan'])
    assert (fdf.loc[bad, 'retval'] == 1).all() and (fdf.loc[~bad, 'retval'] == 0).all()
    assert (fdf['attempts'] == 1).all()
    for _, row in fdf[bad].iterrows():
        assert os.path.exists(row['params']) and os.path.exists(os.path.join(row['outdir'], 'motif_matrix.npy'))
        assert ('at most 32' in open(row['log']).read()) or ('non-finite' in open(row['log']).read())
        assert not os.path.exists(os.path.join(row['outdir'], 'positions_df.tsv'))
    for _, row in fdf[~bad].iterrows():
        assert open(row['log']).read() == ''
        for name in ('positions_df', 'ranks_df'):
            df = pd.read_pickle(os.path.join(row['outdir'], f'{name}.pkl'))
            assert open(os.path.join(row['outdir'], f'{name}.tsv')).read() == df.to_csv(sep='\t', index=False)
    # the good tasks equal a run without the poisoned motifs
    ref = batch_mod.run_batch(ds.replace(), {k: v for k, v in motifs.items() if k.startswith('G')},
                              str(tmp_path / 'ref'), motif_orientations=['+', '+/-'], motif_margin=2)
    for (_, a), (_, b) in zip(fdf[~bad].iterrows(), ref.iterrows()):
        pa, pb = (pd.read_pickle(os.path.join(r['outdir'], 'positions_df.pkl')) for r in (a, b))
        assert pa.equals(pb)


def test_chunked_permutation_staging_equals_the_whole_matrix():
    """A PermutationSource staged in row chunks (Engine._build_y_chunked, how cfg 5's 40 GB index matrix reaches the
    device) gives the same operand, statistics and outputs as the materialised (P, N) matrix -- to the last bit."""
    from mepp_b200 import synth, permutations
    cfg = synth.make_config("cfg1", n_seq=777, n_motifs=4, n_perm=8)
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    P = 200
    src = permutations.PermutationSource(777, P, seed=5)
    whole = _engine().stage(cfg["ascii"], cfg["scores"], src.materialize(), use_gc=True)
    want = whole.profile(tasks, margin=2, return_r=True)
    eng = _engine()
    eng.PERM_STAGE_ROWS = 64                      # four chunks, the last one short
    got = eng.stage(cfg["ascii"], cfg["scores"], src, use_gc=True).profile(tasks, margin=2, return_r=True)
    assert torch.equal(eng._y_rows, whole._y_rows) and torch.equal(eng._col_syz_x, whole._col_syz_x)
    assert got.r.shape[2] == 1 + P and np.allclose(got.r, want.r, rtol=1e-6, atol=1e-9)   # (float64 atomics: order may differ)
    assert np.array_equal(got.density, want.density)


def test_match_driven_sparse_profile_equals_the_entry_form():
    """mepp_profile_matches (one gathered row of Y per match, window sums of the per-position products) against
    mepp_profile_sparse (one per non-zero of the blurred matrix): same r within the float32 rounding of the blurred
    values, for margins 0 .. 16, a peaked profile (planted motif) and a motif that never matches."""
    from mepp_b200 import synth
    cfg = synth.make_config("cfg1", n_seq=700, n_motifs=6, n_perm=60)
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    eng = _engine().stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    for margin in (0, 1, 2, 5, 16):
        eng.sparse_matches = False
        want = eng.profile(tasks, margin=margin, return_r=True)
        eng.sparse_matches = True
        got = eng.profile(tasks, margin=margin, return_r=True)
        assert set(eng.last_paths) == {'sparse'}
        scale = np.abs(want.r).max(axis=-1, keepdims=True)
        assert (np.abs(got.r.astype(np.float64) - want.r) <= 2e-6 * np.abs(want.r) + 2e-6 * scale + 1e-12).all(), margin
        assert np.array_equal(got.density, want.density)
    eng.sparse_matches = False


def test_motif_length_and_margin_extremes():
    """m = 1, 2, 3 and the maximum 32 at the maximum margin 16, on sequences shorter than the blur window
    would like (L = 40) and on an L that is not a multiple of anything (L = 203)."""
    from mepp_b200 import synth
    rng = np.random.default_rng(12)
    for L in (40, 203):
        ascii_u8, scores = synth.synth_sequences(130, L, seed=5 + L, n_rate=0.01, lower_rate=0.01)
        perm = synth.synth_permutations(130, 9, seed=4)
        motifs = []
        for m in (1, 2, 3, 32):
            pwm = rng.dirichlet(np.full(4, 0.05), size=m).T          # sharp columns -> finite thresholds
            motifs.append(pwm)
        eng = _engine().stage(ascii_u8, scores, perm)
        tasks = [(pwm, o) for pwm in motifs for o in ('+', '+/-')]
        batch = eng.profile(tasks, margin=16, return_hits=True, return_r=True)
        codes = staging.ascii_to_codes(ascii_u8)
        Y = staging.permuted_scores(scores, perm)
        z = staging.gc_ratio_global(codes)
        for t, (pwm, o) in enumerate(tasks):
            orc = profile.profile_task(codes, Y, z, pwm, o, margin=16)
            assert np.array_equal(batch.motif_score_matrix(t, L).view(np.uint32), orc["X0"].view(np.uint32)), (L, t)
            assert _r_close(batch.r[t], orc["r"]).all(), (L, t)
            assert np.array_equal(np.asarray(batch.positions["smoothed_count"][t]), orc["positions"]["smoothed_count"])


def test_honour_flags_changes_the_threshold_like_the_oracle():
    """--motifpseudocount / --motifpvalue only reach the scan with honour_flags=True (the reference drops them,
    core.py:513-534); with the flag the matches equal the oracle's at the same pseudocount / p-value."""
    from mepp_b200 import synth
    ascii_u8, scores = synth.synth_sequences(200, 101, seed=21)
    perm = synth.synth_permutations(200, 5, seed=21)
    motifs = synth.synth_motifs(3, seed=8, m_lo=6, m_hi=12)
    eng = _engine().stage(ascii_u8, scores, perm)
    tasks = [(m, '+/-') for m in motifs]
    codes = staging.ascii_to_codes(ascii_u8)
    default = eng.profile(tasks, margin=1, return_hits=True, pseudocount=0.01, pvalue=1e-3)
    honoured = eng.profile(tasks, margin=1, return_hits=True, pseudocount=0.01, pvalue=1e-3, honour_flags=True)
    differs = 0
    for t, (pwm, o) in enumerate(tasks):
        Wf, Wr, thr = profile.motif_weights(pwm, o)
        assert default.thr[t] == thr
        assert np.array_equal(default.motif_score_matrix(t, 101).view(np.uint32),
                              profile.scan(codes, Wf, Wr, thr).view(np.uint32))
        Wf, Wr, thr = profile.motif_weights(pwm, o, pseudocount=0.01, pvalue=1e-3)
        assert honoured.thr[t] == thr
        X0 = honoured.motif_score_matrix(t, 101)
        assert np.array_equal(X0.view(np.uint32), profile.scan(codes, Wf, Wr, thr).view(np.uint32))
        differs += int((X0 != 0).sum() != (default.motif_score_matrix(t, 101) != 0).sum())
    assert differs > 0


def test_heatmap_from_hits_equals_dense_block_max(cfg1):
    from mepp_b200 import heatmap
    eng = _engine().stage(cfg1["ascii"], cfg1["scores"], cfg1["perm_idx"])
    tasks = [(pwm, o) for _, pwm, o in cfg1["tasks"]][:3]
    batch = eng.profile(tasks, margin=cfg1["margin"], return_hits=True)
    for t in range(len(tasks)):
        X0 = batch.motif_score_matrix(t, cfg1["L"])
        for pw, ph in ((100, 50), (201, 1000), (7, 13)):
            a = batch.heatmap(t, cfg1["L"], pw, ph)
            b = heatmap.compress_dense_to_pixels(X0, pw, ph)
            assert a.shape == b.shape and np.array_equal(a, b)



def test_tensor_path_keeps_rows_of_tiny_scores_within_the_gate():
    """Regression for the precision hole of the tiled operand: with ONE global scale (256) the fp16 hi + lo pieces lost
    relative precision on rows whose non-zeros are all tiny (scores barely above the threshold, large margins, few
    sequences: 3e-5 relative in r, found by tools/gpu_fuzz.py).  With the per-task power-of-two scale (mepp_x_scales)
    the randomised small cases stay inside the 1e-5 gate on the tensor path."""
    from mepp_b200 import synth
    rng = np.random.default_rng(20240)
    eng = _engine()
    eng.profile_path = 'tensor'
    worst = 0.0
    for case in range(40):
        N = int(rng.integers(60, 300)); L = int(rng.integers(60, 300)); P = int(rng.choice([7, 33, 100]))
        margin = int(rng.choice([2, 5, 16, 16]))
        ascii_u8, scores = synth.synth_sequences(N, L, seed=int(rng.integers(1 << 30)), n_rate=0.01)
        perm = synth.synth_permutations(N, P, seed=int(rng.integers(1 << 30)))
        tasks = []
        for m in (int(rng.integers(4, 20)) for _ in range(3)):
            pwm = rng.dirichlet(np.full(4, float(rng.choice([0.05, 0.3, 1.0]))), size=m).T
            tasks.append((pwm, str(rng.choice(['+', '-', '+/-']))))
        eng.stage(ascii_u8, scores, perm, use_gc=True)
        b = eng.profile(tasks, margin=margin, return_r=True)
        assert set(eng.last_paths) == {'tensor'}
        codes = staging.ascii_to_codes(ascii_u8)
        Y = staging.permuted_scores(scores, perm)
        z = staging.gc_ratio_global(codes)
        for t, (pwm, o) in enumerate(tasks):
            orc = profile.profile_task(codes, Y, z, pwm, o, margin=margin)
            d = np.abs(b.r[t].astype(np.float64) - orc["r"].astype(np.float64))
            scale = np.abs(orc["r"]).max(axis=-1, keepdims=True)
            bound = 1e-5 * np.abs(orc["r"]) + 1e-5 * scale + 1e-12
            worst = max(worst, float((d / bound).max()))
            assert (d <= bound).all(), (case, t, N, L, margin, float((d / bound).max()))
    assert worst > 0.0


def test_device_heatmap_equals_dense_block_max(cfg1):
    """mepp_heatmap (clip at mean + sigma * std of the non-zero scores, block maximum down to the pixel
    grid, on the device from the match list) against the dense restatement of plot.py:32-66,106-145.  Unclipped pixels
    are bit-equal; clipped ones equal the limit, which the device computes in float64 and numpy in float32."""
    from mepp_b200 import heatmap
    eng = _engine().stage(cfg1["ascii"], cfg1["scores"], cfg1["perm_idx"])
    tasks = [(pwm, o) for _, pwm, o in cfg1["tasks"]][:6]
    ref = eng.profile(tasks, margin=cfg1["margin"], return_hits=True)
    N, L = cfg1["N"], cfg1["L"]
    for pw, ph, sigma in ((100, 50, 1.0), (201, 1000, 1.0), (7, 13, 2.5), (64, 64, None)):
        got = eng.profile(tasks, margin=cfg1["margin"], heatmap=(pw, ph, sigma))
        assert got.heatmaps is not None and got.heatmaps.dtype == np.float32
        assert np.array_equal(np.asarray(got.positions["positional_r"]), np.asarray(ref.positions["positional_r"]))
        for t in range(len(tasks)):
            X0 = ref.motif_score_matrix(t, L)
            want = heatmap.compress_dense_to_pixels(X0, pw, ph, sigma)
            a = got.heatmap(t, L, pw, ph, sigma)
            assert a.shape == want.shape, (pw, ph, t)
            lim = heatmap.clip_limit(X0[X0 != 0], sigma)
            if lim is None:
                assert np.array_equal(a, want)
            else:
                below = want < np.float32(lim) * (1 - 1e-6)
                assert np.array_equal(a[below], want[below]), (pw, ph, t)
                assert np.allclose(a[~below], want[~below], rtol=2e-6, atol=0), (pw, ph, t)
            assert (a > 0).sum() == (want > 0).sum()


def test_shared_scan_units_equal_separate_scans():
    """'+' tasks ride on the '+/-' pass of the same motif (engine.scan_units); out-of-order pairs, unpaired
    tasks and '-' tasks in the same call; results equal the one-scan-per-task path and the oracle."""
    from mepp_b200 import synth
    from mepp_b200.engine import scan_units
    from mepp_b200 import motif_layers
    ascii_u8, scores = synth.synth_sequences(257, 150, seed=31, n_rate=0.02)
    perm = synth.synth_permutations(257, 11, seed=31)
    m = synth.synth_motifs(4, seed=17, m_lo=5, m_hi=20)
    tasks = [(m[0], '+'), (m[0], '+/-'), (m[1], '+/-'), (m[2], '-'), (m[3], '+'), (m[1], '+'), (m[2], '+/-')]
    units = scan_units(motif_layers.scan_tables(tasks))
    assert sorted(map(tuple, units.tolist())) == [(1, 0), (2, 5), (3, -1), (4, -1), (6, -1)]
    eng = _engine().stage(ascii_u8, scores, perm)
    assert eng.share_scans
    shared = eng.profile(tasks, margin=3, return_hits=True, return_r=True)
    assert eng.last_units == 5
    eng.share_scans = False
    separate = eng.profile(tasks, margin=3, return_hits=True, return_r=True)
    assert eng.last_units == 7
    assert np.array_equal(shared.hits, separate.hits)
    assert np.array_equal(shared.density, separate.density)
    assert np.allclose(shared.r, separate.r, rtol=1e-6, atol=1e-9)
    codes = staging.ascii_to_codes(ascii_u8)
    Y = staging.permuted_scores(scores, perm)
    z = staging.gc_ratio_global(codes)
    for t, (pwm, o) in enumerate(tasks):
        orc = profile.profile_task(codes, Y, z, pwm, o, margin=3)
        assert np.array_equal(shared.motif_score_matrix(t, 150).view(np.uint32), orc["X0"].view(np.uint32)), t
        assert _r_close(shared.r[t], orc["r"]).all(), t


def test_scan_is_deterministic_after_longer_motifs():
    """Regression: the pair tables of a scan unit are staged in shared memory in whole ladder buckets (3, 5, 6, 8,
    11, 16 slots); slots past the motif's own must be zero, not what an earlier, longer motif left there."""
    from mepp_b200 import synth
    rng = np.random.default_rng(77)
    ascii_u8, scores = synth.synth_sequences(4096, 512, seed=8, n_rate=0.005)
    perm = synth.synth_permutations(4096, 7, seed=8)
    mk = lambda m: rng.dirichlet(np.full(4, 0.3), size=m).T
    long_tasks = [(mk(32), '+/-') for _ in range(8)]
    short_tasks = [(mk(m), o) for m in (13, 14, 9, 10, 11, 17, 21, 22) for o in ('+', '+/-')]
    eng = _engine().stage(ascii_u8, scores, perm)
    ref = None
    for rep in range(3):
        eng.profile(long_tasks, margin=2)                       # fills every table slot with non-zero entries
        got = eng.profile(short_tasks, margin=2, return_hits=True)
        if ref is None:
            ref = got
            codes = staging.ascii_to_codes(ascii_u8)
            for t in (0, 1, 2, 3):
                Wf, Wr, thr = profile.motif_weights(*short_tasks[t])
                assert np.array_equal(got.motif_score_matrix(t, 512).view(np.uint32),
                                      profile.scan(codes, Wf, Wr, thr).view(np.uint32))
        else:
            assert np.array_equal(got.hits, ref.hits), rep


def test_sparse_and_tensor_profile_paths_agree(monkeypatch):
    """The gather path (entry list -> CSR -> float64 sums of Y rows) and the tcgen05 GEMM path give the same
    profiles; an entry list that overflows falls back to the GEMM path for that group."""
    from mepp_b200 import synth, engine as engine_mod
    ascii_u8, scores = synth.synth_sequences(1500, 260, seed=12, n_rate=0.01)
    perm = synth.synth_permutations(1500, 37, seed=12)
    motifs = synth.synth_motifs(6, seed=4, m_lo=5, m_hi=18)
    tasks = [(m, o) for m in motifs for o in ('+', '+/-')]
    eng = _engine().stage(ascii_u8, scores, perm)
    eng.profile_path = 'sparse'
    sp = eng.profile(tasks, margin=2, return_r=True, return_hits=True)
    assert set(eng.last_paths) == {'sparse'}
    eng.profile_path = 'tensor'
    eng.last_paths = []
    tc = eng.profile(tasks, margin=2, return_r=True, return_hits=True)
    assert set(eng.last_paths) == {'tensor'}
    assert np.array_equal(sp.hits, tc.hits)
    assert _r_close(sp.r, tc.r).all()
    codes = staging.ascii_to_codes(ascii_u8)
    Y = staging.permuted_scores(scores, perm)
    z = staging.gc_ratio_global(codes)
    for t in (0, 1, 6, 7):
        orc = profile.profile_task(codes, Y, z, tasks[t][0], tasks[t][1], margin=2)
        assert _r_close(sp.r[t], orc["r"], rel=2e-6, floor_frac=2e-6).all(), t    # float64 sums: tighter than the GEMM
    for col in ('positional_r', 'permutation_positional_r_lower', 'integral_r'):
        assert np.allclose(np.asarray(sp.positions[col]), np.asarray(tc.positions[col]), rtol=1e-4, atol=1e-7), col
    # a match list that overflows (match-driven form, the default) -> the group is redone from the blurred entries
    eng.profile_path = 'sparse'
    assert eng.sparse_matches
    eng.last_paths, eng.fallbacks = [], []
    fm = eng.profile(tasks, margin=2, return_r=True, hits_cap=64)
    assert eng.last_paths == ['sparse', 'sparse'] and eng.fallbacks == ['entries']
    assert _r_close(fm.r, sp.r, rel=2e-6, floor_frac=2e-6).all()
    # an entry list that overflows (entry form) -> the tensor-core path
    monkeypatch.setattr(engine_mod, 'SPARSE_ENTRY_DENSITY', 1e-9)
    monkeypatch.setattr(engine_mod, 'MIN_ENTRY_CAP', 16)
    eng.sparse_matches = False
    eng.last_paths = []
    fb = eng.profile(tasks, margin=2, return_r=True)
    assert eng.last_paths == ['sparse', 'tensor']
    assert np.array_equal(fb.r, tc.r) or _r_close(fb.r, tc.r).all()


@pytest.mark.parametrize("T,L,P,margin", [(3, 101, 37, 2), (2, 64, 1000, 0), (1, 33, 1, 5), (2, 7, 0, 4)])
def test_device_finalisation_equals_host_finalisation(T, L, P, margin):
    """K4 (mepp_finalize_positions) against stats_host.finalize_into (the numpy restatement of core.py:286-413) on
    the same K3 outputs: interpolated percentiles / p-values / counts identical, float32 Fisher-z columns within
    2 ulp (logf / expf), t-test p-value within 1e-9 relative of scipy's stdtr."""
    import ctypes as C
    from mepp_b200 import _lib, stats_host
    from mepp_b200.engine import _ptr
    _engine()
    rng = np.random.default_rng(T * 1000 + L)
    Cc = 1 + P
    n_seq = 5000
    r_h = (rng.normal(scale=0.02, size=(T * L, Cc))).astype(np.float32)
    r_h[0, 0] = 0.0
    r_h[1, 0] = 0.93                                       # a large |t|: tiny p-value
    r = torch.from_numpy(r_h).cuda()
    count = torch.from_numpy(rng.integers(0, 9, (T, L)).astype(np.int32)).cuda()
    q = stats_host.percentile_q(95)
    dev = [None] * 5
    k_lo = k_hi = 0
    g = [0.0] * 4
    if P > 0:
        k_lo, g[0] = stats_host.virtual_index(P, q[0]); k_hi, g[1] = stats_host.virtual_index(P, q[1])
        ks_lo, g[2] = stats_host.virtual_index(L * P, q[0]); ks_hi, g[3] = stats_host.virtual_index(L * P, q[1])
        dev = [torch.empty((T, L, 4), dtype=torch.float32, device='cuda'), torch.empty((T, L), dtype=torch.int32, device='cuda'),
               torch.empty((T, 4), dtype=torch.float32, device='cuda'), torch.empty((T, L), dtype=torch.int64, device='cuda'),
               torch.empty((T, Cc), dtype=torch.float64, device='cuda')]
        work = torch.empty(int(_lib.lib.mepp_perm_stats_work_bytes(T, L, Cc)), dtype=torch.uint8, device='cuda')
        _lib.check(_lib.lib.mepp_perm_stats(_ptr(r), Cc, T, L, Cc, k_lo, k_hi, ks_lo, ks_hi, *[_ptr(x) for x in dev],
                                            _ptr(work), None), "perm_stats")
    out = torch.empty(int(_lib.lib.mepp_finalize_out_bytes(T, L)), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib.mepp_finalize_positions(_ptr(r), Cc, T, L, Cc, *[_ptr(x) for x in dev], _ptr(count), n_seq, margin,
                                                k_lo, k_hi, g[0], g[1], g[2], g[3],
                                                float(stats_host.z_margin_f32(n_seq, 95)), _ptr(out), None), "finalize")
    torch.cuda.synchronize()
    got = stats_host.alloc_positions(T, L, P)
    stats_host.store_finalized(got, 0, out.cpu().numpy(), T, L)
    ref = stats_host.alloc_positions(T, L, P)
    host = [x.cpu().numpy() if x is not None else None for x in dev]
    stats_host.finalize_into(ref, 0, r_h.reshape(T, L, Cc)[:, :, 0].copy(), *host, count.cpu().numpy(), n_seq, margin, 95)
    for k in ref['pos']:
        a, b = got['pos'][k], ref['pos'][k]
        if k in ('positional_r_lower', 'positional_r_upper'):
            assert np.abs(a - b).max() <= 3e-7 * max(1.0, np.abs(b).max()), k
        elif k == 'positional_r_pval':
            assert np.allclose(a, b, rtol=1e-9, atol=1e-300), (k, np.abs(a / b - 1).max())
        else:
            assert np.array_equal(a, b, equal_nan=True), k
    for k in ref['task']:
        assert np.array_equal(got['task'][k], ref['task'][k]), k


@pytest.mark.parametrize("T,L,P,ties", [(2, 50, 1000, False), (3, 40, 999, False), (2, 33, 257, False),
                                        (1, 20, 2000, False), (2, 30, 1000, True), (2, 21, 100, False)])
def test_perm_stats_against_numpy(T, L, P, ties):
    """K3 on its own: per-row order statistics (tail selection in registers, full sort for heavy ties or wide
    intervals), counts, pooled counts and the pooled order statistics (radix select) against numpy sorts."""
    from mepp_b200 import _lib, stats_host
    from mepp_b200.engine import _ptr
    _engine()
    rng = np.random.default_rng(P + L)
    Cc = 1 + P
    ld = Cc + 3
    r_h = np.zeros((T * L, ld), dtype=np.float32)
    vals = rng.normal(scale=0.03, size=(T * L, Cc)).astype(np.float32)
    if ties:
        vals = (np.round(vals * 40) / 40).astype(np.float32)          # a handful of distinct values
    r_h[:, :Cc] = vals
    r = torch.from_numpy(r_h).cuda()
    for ci in (95, 80):
        q = stats_host.percentile_q(ci)
        k_lo, _ = stats_host.virtual_index(P, q[0]); k_hi, _ = stats_host.virtual_index(P, q[1])
        ks_lo, _ = stats_host.virtual_index(L * P, q[0]); ks_hi, _ = stats_host.virtual_index(L * P, q[1])
        full_os = torch.empty((T, L, 4), dtype=torch.float32, device='cuda'); full_cnt = torch.empty((T, L), dtype=torch.int32, device='cuda')
        semi_os = torch.empty((T, 4), dtype=torch.float32, device='cuda'); semi_cnt = torch.empty((T, L), dtype=torch.int64, device='cuda')
        integral = torch.empty((T, Cc), dtype=torch.float64, device='cuda')
        work = torch.empty(int(_lib.lib.mepp_perm_stats_work_bytes(T, L, Cc)), dtype=torch.uint8, device='cuda')
        _lib.check(_lib.lib.mepp_perm_stats(_ptr(r), ld, T, L, Cc, k_lo, k_hi, ks_lo, ks_hi, _ptr(full_os), _ptr(full_cnt),
                                            _ptr(semi_os), _ptr(semi_cnt), _ptr(integral), _ptr(work), None), "perm_stats")
        torch.cuda.synchronize()
        perm = vals[:, 1:].reshape(T, L, P)
        r0 = vals[:, 0].reshape(T, L)
        srt = np.sort(perm, axis=-1)
        want = np.stack([srt[..., k_lo], srt[..., min(k_lo + 1, P - 1)], srt[..., k_hi], srt[..., min(k_hi + 1, P - 1)]], axis=-1)
        assert np.array_equal(full_os.cpu().numpy(), want), (ci, "per-row order statistics")
        assert np.array_equal(full_cnt.cpu().numpy(), (np.abs(perm) >= np.abs(r0)[..., None]).sum(-1))
        pooled = np.sort(perm.reshape(T, L * P), axis=-1)
        M = L * P
        wants = np.stack([pooled[:, ks_lo], pooled[:, min(ks_lo + 1, M - 1)], pooled[:, ks_hi], pooled[:, min(ks_hi + 1, M - 1)]], axis=-1)
        assert np.array_equal(semi_os.cpu().numpy(), wants), (ci, "pooled order statistics")
        ap = np.abs(perm).reshape(T, 1, L * P)
        assert np.array_equal(semi_cnt.cpu().numpy(), (ap >= np.abs(r0)[..., None]).sum(-1))
        integ = 0.5 * (np.abs(vals.reshape(T, L, Cc)[:, 1:].astype(np.float64)) + np.abs(vals.reshape(T, L, Cc)[:, :-1].astype(np.float64)))
        assert np.allclose(integral.cpu().numpy(), integ.sum(1), rtol=1e-12)


def test_device_thresholds_equal_host_thresholds():
    """The threshold DP on the device (mepp_thresholds_dp) gives bit-identical thresholds, biases and pass bounds
    to the host DP (mepp_motif_table = MOODS threshold_from_p) for every motif of the shipped libraries, all
    three orientations, and for honoured --motifpseudocount / --motifpvalue / background."""
    from mepp_b200 import synth, motif_layers
    eng = _engine()
    tasks = []
    for lib in ("ohler", "homer"):
        motifs = synth.load_motif_library(lib)
        tasks += [(pwm, o) for pwm in list(motifs.values())[:80] for o in ('+', '-', '+/-')]
    for kw in (dict(), dict(honour_flags=True, pseudocount=0.01, pvalue=1e-3, bg=[0.3, 0.2, 0.2, 0.3])):
        dt = eng.device_tables_for(tasks, **kw)
        ref = motif_layers.scan_tables(tasks, **kw)
        thr = dt.dev['thr'].cpu().numpy()
        bias = dt.dev['bias'].cpu().numpy()
        qthr = dt.dev['qthr'].cpu().numpy().view(np.uint32)
        assert np.array_equal(thr, np.array([tb['thr'] for tb in ref]))
        assert np.array_equal(bias, np.array([tb['bias'] for tb in ref], dtype=np.float32))
        assert np.array_equal(qthr, np.stack([tb['qthr'] for tb in ref]))


@pytest.mark.parametrize("grid", [1000.0, 2000.0])
def test_device_thresholds_equal_the_oracle_on_both_dp_grids(grid):
    """MOODS threshold_from_p on the 1/1000 grid (MOODS <= 1.9.2) and on the 1/2000 grid (MOODS >= 1.9.3, the
    default): the device DP, the host C++ DP and the numpy oracle give the same threshold for EVERY motif of the three
    shipped libraries in the orientations that have their own threshold ('+' and '-'), and a task run with
    dp_precision=grid reports exactly that threshold."""
    from mepp_b200 import synth, motif_layers
    from oracle import moods
    eng = _engine()
    tasks, want = [], []
    for lib in ("ohler", "homer_clustered", "homer"):
        for pwm in synth.load_motif_library(lib).values():
            W = moods.log_odds(pwm, [0.25] * 4, 1e-4)
            for o, Wo in (('+', W), ('-', W[::-1, ::-1])):
                tasks.append((pwm, o))
                want.append(moods.threshold_from_p(Wo, [0.25] * 4, 1e-4, grid))
    dt = eng.device_tables_for(tasks, dp_precision=grid)
    host = motif_layers.scan_tables(tasks, dp_precision=grid)
    assert np.array_equal(dt.dev['thr'].cpu().numpy(), np.array(want))
    assert np.array_equal(np.array([tb['thr'] for tb in host]), np.array(want))
    cfg = synth.make_config("cfg1", n_seq=64, n_motifs=3, n_perm=8)
    sel = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    got = eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"]).profile(sel, margin=2, dp_precision=grid)
    for (pwm, o), thr in zip(sel, got.thr):
        assert thr == profile.motif_weights(pwm, o, precision=grid)[2]


def test_tiled_operand_is_restored_even_when_the_match_list_overflows():
    """Tensor path: the scan stores only non-zero pieces into an all-zero operand and mepp_x_planes_reset clears them
    from the match list; with a match list that is too small the whole operand is cleared instead.  Either way a
    later call on the same engine (other tasks, other margin) must see a clean operand."""
    from mepp_b200 import synth
    ascii_u8, scores = synth.synth_sequences(700, 180, seed=44, n_rate=0.01)
    perm = synth.synth_permutations(700, 9, seed=44)
    m = synth.synth_motifs(6, seed=23, m_lo=5, m_hi=14)
    first = [(m[0], '+/-'), (m[1], '+'), (m[2], '-')]
    second = [(m[3], '+'), (m[4], '+/-'), (m[5], '+/-'), (m[0], '+')]
    ref_eng = _engine().stage(ascii_u8, scores, perm)
    ref_eng.profile_path = 'tensor'
    want = ref_eng.profile(second, margin=1, return_r=True)
    for cap in (None, 4):                                     # list large enough / overflowing
        eng = _engine().stage(ascii_u8, scores, perm)
        eng.profile_path = 'tensor'
        eng.profile(first, margin=5, hits_cap=cap)
        got = eng.profile(second, margin=1, return_r=True)
        assert np.allclose(got.r, want.r, rtol=1e-6, atol=1e-9), cap      # (row sums are float64 atomics: order may differ)
        assert np.array_equal(got.density, want.density)


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_device_statistics_equal_the_reference_golden_vectors(name):
    """K3 + K4 (mepp_perm_stats, mepp_finalize_positions) on the r matrices of tests/golden/reference_stats.npz
    against the columns the reference's own get_positional_r_df / add_parametric_confidence_intervals_to_positional_r_df
    produced for them (tests/golden/make_reference_stat_fixtures.py)."""
    import os
    from mepp_b200 import _lib, stats_host
    from mepp_b200.engine import _ptr
    _engine()
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_stats.npz")) as z:
        ref = {k: z[k] for k in z.files}
    L, P, center, ci, n = (int(v) for v in ref[f"stat_{name}__args"])
    want = {str(c): ref[f"stat_{name}__col__{c}"] for c in ref[f"stat_{name}__columns"]}
    r_h = ref[f"stat_{name}__r"]
    r_h = np.ascontiguousarray(r_h if r_h.ndim > 1 else r_h[:, None])
    Cc = 1 + P
    r = torch.from_numpy(r_h).cuda()
    count = torch.zeros((1, L), dtype=torch.int32, device='cuda')
    q = stats_host.percentile_q(ci)
    dev, k_lo, k_hi, g = [None] * 5, 0, 0, [0.0] * 4
    if P > 0:
        k_lo, g[0] = stats_host.virtual_index(P, q[0]); k_hi, g[1] = stats_host.virtual_index(P, q[1])
        ks_lo, g[2] = stats_host.virtual_index(L * P, q[0]); ks_hi, g[3] = stats_host.virtual_index(L * P, q[1])
        dev = [torch.empty((1, L, 4), dtype=torch.float32, device='cuda'), torch.empty((1, L), dtype=torch.int32, device='cuda'),
               torch.empty((1, 4), dtype=torch.float32, device='cuda'), torch.empty((1, L), dtype=torch.int64, device='cuda'),
               torch.empty((1, Cc), dtype=torch.float64, device='cuda')]
        work = torch.empty(int(_lib.lib.mepp_perm_stats_work_bytes(1, L, Cc)), dtype=torch.uint8, device='cuda')
        _lib.check(_lib.lib.mepp_perm_stats(_ptr(r), Cc, 1, L, Cc, k_lo, k_hi, ks_lo, ks_hi, *[_ptr(x) for x in dev],
                                            _ptr(work), None), "perm_stats")
    out = torch.empty(int(_lib.lib.mepp_finalize_out_bytes(1, L)), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib.mepp_finalize_positions(_ptr(r), Cc, 1, L, Cc, *[_ptr(x) for x in dev], _ptr(count), n, 0,
                                                k_lo, k_hi, g[0], g[1], g[2], g[3],
                                                float(stats_host.z_margin_f32(n, ci)), _ptr(out), None), "finalize")
    torch.cuda.synchronize()
    store = stats_host.alloc_positions(1, L, P)
    stats_host.store_finalized(store, 0, out.cpu().numpy(), 1, L)
    got = stats_host.positions_view(store, None if center < 0 else center)
    for k in want:
        a, b = np.asarray(got[k][0], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        if k in ('positional_r_lower', 'positional_r_upper'):
            assert np.abs(a - b).max() <= 4e-7, k
        elif k == 'positional_r_pval':
            assert np.allclose(a, b, rtol=2e-5, atol=1e-300), k
        elif k in ('integral_r', 'permutation_integral_r_lower', 'permutation_integral_r_upper'):
            assert np.allclose(a, b, rtol=3e-7), k                    # float32 trapezoid sums: summation order
        else:
            assert np.array_equal(a, b), k


def test_bit_planes_match_the_codes(cfg1):
    """mepp_pack_planes: "a letter other than c" planes of the staged sequences (N, positions past L and padding
    sequences are zero in every plane)."""
    eng = _engine()
    eng.bitpar = True
    eng.stage(cfg1["ascii"], cfg1["scores"], cfg1["perm_idx"])
    N, L = cfg1["N"], cfg1["L"]
    codes = staging.ascii_to_codes(cfg1["ascii"])
    pl = eng.planes_T.cpu().numpy().view(np.uint32)               # (WB, 4, n_pad)
    WB = pl.shape[0]
    assert (pl[0] == 0).all() and (pl[:, :, N:] == 0).all(), (N, pl.shape)
    bits = ((pl[1:, :, :N, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8)   # (WB-1, 4, N, 32)
    bits = bits.transpose(1, 2, 0, 3).reshape(4, N, (WB - 1) * 32)
    for c in range(4):
        want = np.zeros((N, (WB - 1) * 32), dtype=np.uint8)
        want[:, :L] = (codes != c) & (codes < 4)
        assert np.array_equal(bits[c], want), c


@pytest.mark.parametrize("cut", [0.0, 1e9])
@pytest.mark.parametrize("orientations", [('+', '+/-'), ('-',), ('+/-',)])
def test_scan_first_stages_agree_with_the_oracle(cut, orientations):
    """Both first stages of the scan -- integer pair tables (cost ratio 0: only never-matching units take the
    bit-parallel kernel) and bit-parallel bad-letter counters (huge ratio: every unit with a sound (d0, K)) -- feed the
    same exact evaluation: matches bit-identical to the oracle; motif lengths 1..32, N-rich ragged sequences,
    the shipped ohler motifs, a never-matching uniform motif, margins 0 and 7."""
    from mepp_b200 import synth
    ascii_u8, scores = synth.synth_sequences(203, 171, seed=41, n_rate=0.03, lower_rate=0.01)
    ascii_u8[5, :] = ord('N')
    ascii_u8[6, 3:] = ord('A')
    perm = synth.synth_permutations(203, 9, seed=41)
    rng = np.random.default_rng(12)
    motifs = [rng.dirichlet(np.full(4, a), size=m).T for m, a in ((1, 0.3), (2, 0.3), (5, 0.2), (8, 0.3), (12, 0.5),
                                                                   (17, 0.3), (25, 1.0), (32, 0.3))]
    motifs += list(synth.load_motif_library("ohler").values())[:4] + [np.full((4, 7), 0.25)]
    cons = np.full((4, 9), 1e-3); cons[[0, 0, 0, 0, 0, 0, 0, 0, 0], np.arange(9)] = 1.0      # poly-A: row 6 matches
    motifs.append(cons / cons.sum(axis=0))
    tasks = [(m, o) for m in motifs for o in orientations]
    codes = staging.ascii_to_codes(ascii_u8)
    for margin in (0, 7):
        eng = _engine()
        eng.bitpar = True
        eng.bp_cost_ratio = cut
        eng.stage(ascii_u8, scores, perm)
        for use_device_tables in (True, False):
            eng.device_tables = use_device_tables
            got = eng.profile(tasks, margin=margin, return_hits=True)
            bp, _scratch, U = eng.last_bp
            use = ((bp.cpu().numpy().view(np.uint32)[:, 64] >> 8) & 1).astype(bool)
            if cut == 0.0:
                assert use.sum() <= 5 * len(orientations)           # only units whose threshold is out of reach
            else:
                assert use.sum() >= len(tasks) // 2
            n_hits = 0
            for t, (pwm, o) in enumerate(tasks):
                Wf, Wr, thr = profile.motif_weights(pwm, o)
                X0 = profile.scan(codes, Wf, Wr, thr)
                n_hits += int((X0 > 0).sum())
                assert np.array_equal(got.motif_score_matrix(t, 171).view(np.uint32), X0.view(np.uint32)), (t, o, margin)
            assert n_hits > 15


def test_full_size_config2_properties():
    """BASELINE.json configs[1] at full size (20 000 x 1 kb, 162 HOMER-clustered motifs x {+, +/-}, 1000 permutations,
    margin 2) through size-independent properties: (1) checksum of checksums -- per-position counts, per-sequence
    densities and the match counter of every task add up to the same number; (2) the matches of a sample of sequences
    (a sequence's matches do not depend on the other sequences) are bit-identical to the oracle's scan of that sample;
    (3) a '+/-' task matches wherever its '+' twin does, with a score at least as high; (4) the overlapped schedule,
    the serial schedule and the bit-parallel first stage give identical outputs; (5) |r| <= 1 and the permutation
    p-values are multiples of 1 / P inside [0, 1]."""
    from mepp_b200 import synth
    cfg = synth.make_config("cfg2")
    N, L, P = cfg["N"], cfg["L"], cfg["P"]
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    eng = _engine().stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"])
    assert eng.overlap
    got = eng.profile(tasks, margin=cfg["margin"])
    count = np.asarray(got.positions["count"])
    assert np.array_equal(count.sum(axis=1), got.density.sum(axis=1))
    assert int(got.hit_counts.sum()) == int(count.sum()) > 100 * len(tasks)      # (match counters are per task group)
    r0 = np.asarray(got.positions["positional_r"])
    assert np.isfinite(r0).all() and (np.abs(r0) <= 1.0 + 1e-6).all()
    for col in ("permutation_full_positional_r_pval", "permutation_semi_positional_r_pval", "permutation_integral_r_pval"):
        p = np.asarray(got.positions[col])
        assert (p >= 0).all() and (p <= 1).all()
    pf = np.asarray(got.positions["permutation_full_positional_r_pval"]) * P
    assert np.abs(pf - np.round(pf)).max() < 1e-6
    # (4) schedules and first stages
    eng.overlap = False
    serial = eng.profile(tasks, margin=cfg["margin"])
    for k in got.positions:
        assert np.array_equal(np.asarray(got.positions[k]), np.asarray(serial.positions[k]), equal_nan=True), k
    assert np.array_equal(got.density, serial.density)
    eng2 = _engine()
    eng2.bitpar = True
    eng2.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"])
    bitpar = eng2.profile(tasks, margin=cfg["margin"])
    assert np.array_equal(got.density, bitpar.density)
    assert np.array_equal(count, np.asarray(bitpar.positions["count"]))
    assert np.array_equal(r0, np.asarray(bitpar.positions["positional_r"]))
    # (2) a sample of sequences against the oracle, a sample of tasks (every 23rd motif, both orientations)
    sample = np.r_[0:40, N // 2:N // 2 + 40, N - 40:N]
    pick = [t for t in range(len(tasks)) if (t // 2) % 23 == 0]
    sub = eng.profile([tasks[t] for t in pick], margin=cfg["margin"], return_hits=True)
    codes = staging.ascii_to_codes(cfg["ascii"][sample])
    for j, t in enumerate(pick):
        Wf, Wr, thr = profile.motif_weights(*tasks[t])
        X0 = sub.motif_score_matrix(j, L)
        assert np.array_equal(X0[sample].view(np.uint32), profile.scan(codes, Wf, Wr, thr).view(np.uint32)), t
        assert np.array_equal((X0 > 0).sum(axis=1), got.density[t])
    # (3) '+' (even rows) against '+/-' (odd rows) of the same motif
    for j in range(0, len(pick) - 1, 2):
        assert tasks[pick[j]][1] == '+' and tasks[pick[j + 1]][1] == '+/-'
        Xp, Xd = sub.motif_score_matrix(j, L), sub.motif_score_matrix(j + 1, L)
        assert (Xd >= Xp).all()


def test_bitpar_records_are_sound_for_the_shipped_libraries():
    """The (cut, count) choice of the bit-parallel first stage is made on the device (bp_tables_kernel).  Independent
    check in numpy, for every motif of the three shipped libraries and the orientations '+', '-', '+/-': with the
    counted taps and good letters read back from the task records, the sum of the K smallest per-tap minima of the
    bad letters' penalties exceeds  max score - threshold + rounding slack  for each filter -- a window the counters
    reject cannot match (scan.cu, "bit-parallel first stage")."""
    from mepp_b200 import synth, motif_layers, _lib
    eng = _engine()
    eng.bitpar = True
    eng.bp_cost_ratio = 1e9
    tasks = []
    for name in ("ohler", "homer_clustered", "homer"):
        for pwm in synth.load_motif_library(name).values():
            tasks += [(pwm, o) for o in ('+', '-', '+/-')]
    ascii_u8, scores = synth.synth_sequences(64, 80, seed=1)
    eng.stage(ascii_u8, scores, None)
    tables = motif_layers.scan_tables(tasks)
    dt = eng.upload_tables(tables)
    bp = dt.dev['bp'].cpu().numpy().view(np.uint32)
    assert bp.shape == (len(tasks), _lib.BP_WORDS)
    n_used = n_never = 0
    for t, tb in enumerate(tables):
        m, nf = tb['m'], tb['n_filters']
        K, use, nfr = int(bp[t, 64] & 0xff), int((bp[t, 64] >> 8) & 1), int((bp[t, 64] >> 16) & 0xff)
        assert nfr == nf
        if not use:
            continue
        counts = [(int(bp[t, 65] >> (16 * f)) & 0xff, int(bp[t, 65] >> (16 * f + 8)) & 0xff) for f in range(2)]
        for f in range(nf):
            W = tb['lut'][:m, :, f].astype(np.float64)                       # (m, 4) float32 weights as doubles
            pen = W.max(axis=1, keepdims=True) - W
            D = W.max(axis=1).sum() + float(tb['bias']) + 2.0 * (4.0 * m) * np.abs(W).max(axis=1).sum() * 5.97e-8 + 1e-6
            if K == 0:
                assert D <= 0.0, (t, f)                                      # nothing can match
                n_never += f == 0
                continue
            n1, n2 = counts[f]
            mins, taps = [], set()
            for i in range(n1 + n2):
                w = int(bp[t, 32 * f + i])
                j, good = w & 0xff, {(w >> 8) & 0xff} | ({(w >> 16) & 0xff} if i >= n1 else set())
                assert j < m and j not in taps and all(g < 4 for g in good), (t, f, i)
                taps.add(j)
                bad = [pen[j, c] for c in range(4) if c not in good]
                mins.append(min(bad))
                assert min(bad) > max(pen[j, c] for c in good), (t, f, j)   # every bad letter costs more than a good one
            assert len(mins) >= K and np.sort(mins)[:K].sum() > D, (t, f, K)
        n_used += 1
    assert n_used > len(tasks) // 2 and n_never > 0


@pytest.mark.parametrize("N,L,margin,seed,long_run", [(333, 77, 0, 1, True), (333, 150, 2, 2, False),
                                                      (1000, 201, 5, 3, False), (257, 64, 16, 4, False),
                                                      (1000, 201, 2, 7, 'many')])
def test_lean_scan_equals_full_scan(N, L, margin, seed, long_run):
    """mepp_scan_lean (matches only + entries built from per-(task, k-block) buckets, the default on the sparse path)
    against mepp_scan (blur and entries from staging rows): identical matches, counts and densities, r within 1e-6;
    runs of overlapping matches in one sequence; a long run overflows its bucket: behind the CUDA-core lean scan the
    group falls back to the full-form scan on the sparse path, behind the tensor-core scan only the overflowed bucket is
    redone (overflow list + bucket_overflow_kernel).  The lean form is run with both first stages: the CUDA-core pair
    tables (mepp_scan_lean) and the tcgen05 int8 products over the one-hot stream (mepp_scan_tc, the default)."""
    from mepp_b200 import synth
    ascii_u8, scores = synth.synth_sequences(N, L, seed=seed, n_rate=0.02, lower_rate=0.01)
    if long_run == 'many':
        ascii_u8[::9, :] = ord('A')                      # every bucket of 32 sequences holds several hundred matches
    elif long_run:
        ascii_u8[6, :] = ord('A')
    else:
        ascii_u8[6, 20:32] = ord('A')
    perm = synth.synth_permutations(N, 13, seed=seed)
    motifs = synth.synth_motifs(6, seed=seed, m_lo=4, m_hi=min(32, L // 3))
    cons = np.full((4, 9), 1e-3); cons[0, :] = 1.0
    motifs.append(cons / cons.sum(axis=0))
    tasks = [(m, o) for m in motifs for o in ('+', '+/-')] + [(motifs[0], '-')]
    res, paths, scans, fallbacks = [], [], [], []
    for lean, tc in ((False, False), (True, False), (True, True)):
        eng = _engine()
        eng.lean, eng.tc_scan = lean, tc
        eng.stage(ascii_u8, scores, perm)
        res.append(eng.profile(tasks, margin=margin, return_hits=True, return_r=True))
        paths.append(list(eng.last_paths))
        scans.append(eng.last_scan)
        fallbacks.append(list(eng.fallbacks))
    full = res[0]
    # a bucket overflow is redone by the full-form scan, still on the sparse path (no detour over the tiled operands)
    assert paths[0] == ['sparse'] and fallbacks[0] == []
    assert paths[1] == (['sparse', 'sparse'] if long_run else ['sparse']), paths
    assert fallbacks[1] == (['full-scan'] if long_run else []), fallbacks
    assert paths[2] == ['sparse'] and fallbacks[2] == [], (paths, fallbacks)      # per-bucket redo, no group fallback
    assert scans == (['full', 'full', 'tc'] if long_run else ['full', 'lean', 'tc']), scans
    assert len(full.hits) > 20
    for lean in res[1:]:
        assert np.array_equal(full.hits, lean.hits)
        assert np.array_equal(full.density, lean.density)
        assert np.array_equal(np.asarray(full.positions['count']), np.asarray(lean.positions['count']))
        assert np.allclose(full.r, lean.r, rtol=1e-6, atol=1e-9)
    if not long_run:
        assert np.array_equal(res[1].r, res[2].r)      # the two lean forms publish the same matches: identical entries
    lean = res[2]
    codes = staging.ascii_to_codes(ascii_u8)
    for t in (0, 1, len(tasks) - 3, len(tasks) - 2):
        Wf, Wr, thr = profile.motif_weights(*tasks[t])
        assert np.array_equal(lean.motif_score_matrix(t, L).view(np.uint32), profile.scan(codes, Wf, Wr, thr).view(np.uint32)), t

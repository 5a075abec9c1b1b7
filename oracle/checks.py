"""Accuracy gates of SURVEY.md section 8(d), computed by the oracle for a few tasks of a workload at FULL size.

TEST INFRASTRUCTURE ONLY (tests/ and bench.py's accuracy block): the device results of the selected tasks are
compared with the oracle's restatement of mepp/motif_layers.py:12-165 (scan, declared float32 order),
mepp/core.py:59-145 (blur), core.py:182-267 (positional / partial correlation, here in float64) and
core.py:269-376 (permutation statistics).  The correlation uses a sparse product (X holds ~1e-3 non-zeros), which
is the same sum as the dense one of oracle/profile.py with the zero terms left out.
"""
import os

import numpy as np
import scipy.sparse as sp

from . import cref, profile


def _r_ab(sa, sb, saa, sbb, sab, n):
    with np.errstate(invalid='ignore', divide='ignore'):
        return ((n * sab) - (sa * sb)) / np.sqrt(((n * saa) - (sa * sa)) * ((n * sbb) - (sb * sb)))


def correlation_sparse(X32, Y64, z64, ystats=None):
    """core.py:195-267 in float64 for a mostly-zero X (N, L) float32."""
    n = X32.shape[0]
    Xs = sp.csr_matrix(X32.T.astype(np.float64))          # (L, N)
    sx = np.asarray(Xs.sum(axis=1))                        # (L, 1)
    sxx = np.asarray(Xs.multiply(Xs).sum(axis=1))
    sxy = Xs @ Y64                                         # (L, C)
    if ystats is None:
        ystats = column_stats(Y64, z64)
    sy, syy, syz, sz, szz = ystats
    r_xy = _r_ab(sx, sy, sxx, syy, sxy, n)
    if z64 is None:
        return r_xy
    sxz = (Xs @ z64)[:, None]
    r_xz = _r_ab(sx, sz, sxx, szz, sxz, n)
    r_yz = _r_ab(sy, sz, syy, szz, syz, n)
    with np.errstate(invalid='ignore', divide='ignore'):
        return (r_xy - r_xz * r_yz) / np.sqrt((1 - r_xz * r_xz) * (1 - r_yz * r_yz))


def column_stats(Y64, z64):
    sy = Y64.sum(0)[None, :]
    syy = (Y64 * Y64).sum(0)[None, :]
    if z64 is None:
        return sy, syy, None, None, None
    return sy, syy, (Y64 * z64[:, None]).sum(0)[None, :], z64.sum(), (z64 * z64).sum()


def borderline_count(codes, Wf32, Wr32, thr, width=1e-5):
    """Windows whose float64 score lies within `width` of the threshold (the only ones whose match status can
    depend on the float32 summation order)."""
    N, L = codes.shape
    n = 0
    for W in (Wf32, Wr32):
        if W is None:
            continue
        W64 = W.astype(np.float64)
        m = W64.shape[1]
        mean = W64.mean(axis=0)
        acc = np.zeros((N, L - m + 1), dtype=np.float64)
        for j in range(m):
            c = codes[:, j:j + L - m + 1]
            acc += np.where(c >= 4, mean[j], W64[np.minimum(c, 3), j])
        n += int((np.abs(acc - thr) < width).sum())
    return n


def accuracy_report(codes, Y32, z32, tasks, task_ids, margin, gpu_hits, gpu_r, gpu_positions, confidence_interval_pct=95,
                    threads=None, borderline_sample=2000):
    """tasks: [(pwm, orientation)] of the selected tasks; gpu_hits: structured (task, seq, pos, x0) with task =
    index into `tasks`; gpu_r: (len(tasks), L, C) float32; gpu_positions: {column: (len(tasks), L)}.
    Returns the dict that bench.py prints as `accuracy`."""
    threads = threads or os.cpu_count() or 1
    N, L = codes.shape
    Y64 = Y32.astype(np.float64)
    z64 = None if z32 is None else z32.astype(np.float64)
    ystats = column_stats(Y64, z64)
    P = Y32.shape[1] - 1
    out = dict(tasks_checked=[int(t) for t in task_ids], sequences_checked=int(N), matches_checked=0,
               match_sets_equal=True, match_scores_bit_equal=True, borderline_windows_f64=0,
               borderline_sample=f"first {min(borderline_sample, N)} sequences, |s64 - thr| < 1e-5",
               r_max_abs_err=0.0, r_max_rel_err=0.0, r_within_1e5_gate=True,
               r_gate="|dr| <= 1e-5 |r| + 1e-5 max_c |r[l, c]| against the float64 oracle",
               ci_max_abs_err=0.0, pvalue_flips=0, pvalue_max_abs_diff=0.0, pvalue_entries=0)
    for k, (pwm, orient) in enumerate(tasks):
        Wf, Wr, thr = profile.motif_weights(pwm, orient)
        X0 = cref.scan(codes, Wf, Wr, thr, threads=threads)
        seq, pos = np.nonzero(X0)
        h = gpu_hits[gpu_hits['task'] == k]
        order = np.lexsort((h['pos'], h['seq']))
        h = h[order]
        out['matches_checked'] += int(seq.size)
        same = h.size == seq.size and np.array_equal(h['seq'], seq) and np.array_equal(h['pos'], pos)
        out['match_sets_equal'] &= bool(same)
        if same:
            out['match_scores_bit_equal'] &= bool(np.array_equal(h['x0'].view(np.uint32), X0[seq, pos].view(np.uint32)))
        else:
            out['match_scores_bit_equal'] = False
        ns = min(borderline_sample, N)
        out['borderline_windows_f64'] += borderline_count(codes[:ns], Wf, Wr, thr)
        X = cref.blur(X0, margin, threads=threads) if margin > 0 else X0
        r64 = correlation_sparse(X, Y64, z64, ystats)
        r64 = np.nan_to_num(r64, nan=0.0)
        rg = gpu_r[k].astype(np.float64)
        d = np.abs(rg - r64)
        scale = np.abs(r64).max(axis=1, keepdims=True)
        out['r_max_abs_err'] = max(out['r_max_abs_err'], float(d.max()))
        with np.errstate(invalid='ignore', divide='ignore'):
            rel = np.where(scale > 0, d / (np.abs(r64) + scale), 0.0)
        out['r_max_rel_err'] = max(out['r_max_rel_err'], float(rel.max()))
        out['r_within_1e5_gate'] &= bool((d <= 1e-5 * np.abs(r64) + 1e-5 * scale + 1e-12).all())
        if P > 0 and gpu_positions is not None:
            cols = profile.positional_r_columns(np.nan_to_num(r64.astype(np.float32), 0.0),
                                                confidence_interval_pct=confidence_interval_pct)
            for name in ('permutation_full_positional_r_pval', 'permutation_semi_positional_r_pval',
                         'permutation_integral_r_pval'):
                a, b = np.asarray(gpu_positions[name][k], dtype=np.float64), np.asarray(cols[name], dtype=np.float64)
                if name == 'permutation_integral_r_pval':
                    a, b = a[:1], b[:1]                        # one value per task, broadcast over the rows
                denom = float(P) if 'semi' not in name else float(P) * L
                out['pvalue_flips'] += int(np.rint(np.abs(a - b) * denom).sum())
                out['pvalue_max_abs_diff'] = max(out['pvalue_max_abs_diff'], float(np.abs(a - b).max()))
                out['pvalue_entries'] += int(a.size)
            for name in ('permutation_full_positional_r_lower', 'permutation_full_positional_r_upper',
                         'permutation_semi_positional_r_lower', 'permutation_semi_positional_r_upper'):
                a, b = np.asarray(gpu_positions[name][k], dtype=np.float64), np.asarray(cols[name], dtype=np.float64)
                out['ci_max_abs_err'] = max(out['ci_max_abs_err'], float(np.abs(a - b).max()))
    out['pvalue_flip_note'] = "flips = comparisons |r_p| >= |r_0| decided differently (p = count / P, or / (L P) pooled)"
    return out

"""Per-motif summary statistics of the profiles.

Restates mepp/utils.py:391-530 of the reference (get_minmax_stats, get_minmax_stats_df) on the
engine's results, so that the tabular part of results_table_orientation_*.{tsv,pkl} can be written
without statsmodels / matplotlib (images, dendrograms and HTML stay out of scope).
`multipletests` restates the statsmodels.stats.multitest procedures the `--mtmethod` flag offers
(statsmodels is not installed here: PARITY UNPINNED for this module; pinned by textbook
known-answer tests in tests/test_host.py).
"""
import numpy as np
import pandas as pd


def _fdrcorrection_sorted(p, alpha, by=False):
    n = p.size
    ecdf = np.arange(1, n + 1) / float(n)
    if by:
        ecdf = ecdf / np.sum(1.0 / np.arange(1, n + 1))
    reject = p <= ecdf * alpha
    if reject.any():
        reject[:np.max(np.nonzero(reject)[0])] = True
    corr = np.minimum.accumulate((p / ecdf)[::-1])[::-1]
    return reject, np.minimum(corr, 1.0)


def multipletests(pvals, alpha=0.05, method='fdr_tsbky'):
    """(reject, pvals_corrected) like statsmodels.stats.multitest.multipletests(...)[:2]."""
    pvals = np.asarray(pvals, dtype=np.float64)
    n = pvals.size
    if n == 0:
        return np.zeros(0, dtype=bool), np.zeros(0)
    order = np.argsort(pvals, kind='stable')
    p = pvals[order]
    method = method.lower()
    reject = None
    if method in ('bonferroni', 'b'):
        corr = p * n
    elif method in ('sidak', 's'):
        corr = -np.expm1(n * np.log1p(-p))
    elif method in ('holm-sidak', 'hs'):
        corr = np.maximum.accumulate(-np.expm1(np.arange(n, 0, -1) * np.log1p(-p)))
    elif method in ('holm', 'h'):
        corr = np.maximum.accumulate(p * np.arange(n, 0, -1))
    elif method in ('simes-hochberg', 'sh'):
        corr = np.minimum.accumulate((p * np.arange(n, 0, -1))[::-1])[::-1]
    elif method in ('hommel', 'ho'):
        a = p.copy()
        for m in range(n, 1, -1):
            cim = np.min(m * p[-m:] / np.arange(1, m + 1.0))
            a[-m:] = np.maximum(a[-m:], cim)
            a[:-m] = np.maximum(a[:-m], np.minimum(m * p[:-m], cim))
        corr = a
    elif method in ('fdr_bh', 'fdr_i', 'fdr_p', 'fdri', 'fdrp'):
        reject, corr = _fdrcorrection_sorted(p, alpha)
    elif method in ('fdr_by', 'fdr_n', 'fdr_c', 'fdrn', 'fdrcorr'):
        reject, corr = _fdrcorrection_sorted(p, alpha, by=True)
    elif method in ('fdr_tsbky', 'fdr_2sbky', 'fdr_twostage', 'fdr_tsbh', 'fdr_2sbh'):
        bky = method in ('fdr_tsbky', 'fdr_2sbky', 'fdr_twostage')
        fact = 1.0 + alpha if bky else 1.0
        alpha_prime = alpha / fact
        reject, corr = _fdrcorrection_sorted(p, alpha_prime)
        r1 = int(reject.sum())
        if r1 == 0 or r1 == n:
            corr = corr * fact
        else:
            n0 = float(n - r1)                      # maxiter = 1: one second-stage pass
            reject, corr = _fdrcorrection_sorted(p, alpha_prime * n / n0)
            corr = corr * n0 / n
            if bky:
                corr = corr * (1.0 + alpha)
    else:
        raise ValueError(f'unknown multiple-testing method {method!r}')
    corr = np.minimum(corr, 1.0)
    if reject is None:
        reject = corr <= alpha
    out_c = np.empty(n)
    out_r = np.empty(n, dtype=bool)
    out_c[order] = corr
    out_r[order] = reject
    return out_r, out_c


def get_minmax_stats(positional_r_df):
    """utils.py:391-446 for one task's positions_df."""
    df = positional_r_df
    cols = list(df.columns)
    has_perm = 'permutation_positional_r_lower' in cols and 'permutation_positional_r_upper' in cols
    prefix = 'permutation_' if has_perm else ''
    r = df['positional_r'].to_numpy()
    imin, imax = int(np.argmin(r)), int(np.argmax(r))      # first occurrence, like df[...].head(1)
    pval = df[f'{prefix}positional_r_pval'].to_numpy()
    pos = df['position'].to_numpy()
    # list(df[col])[0] in the reference yields Python floats: everything below is float64 arithmetic (utils.py:401-444)
    min_r, max_r = float(r[imin]), float(r[imax])
    pval, pos = pval.tolist(), pos.tolist()
    if np.abs(max_r) >= np.abs(min_r):
        ext = (max_r, pval[imax], pos[imax])
    else:
        ext = (min_r, pval[imin], pos[imin])
    if 'integral_r' in cols:
        integ = (df['integral_r'].mean(), df['permutation_integral_r_lower'].mean(),
                 df['permutation_integral_r_upper'].mean(), df['permutation_integral_r_pval'].mean())
    else:
        integ = (None, None, None, None)
    return dict(integral_r=integ[0], integral_r_lower=integ[1], integral_r_upper=integ[2], integral_r_pval=integ[3],
                extreme_r=ext[0], abs_extreme_r=np.abs(ext[0]), extreme_r_pval=ext[1], extreme_r_pos=ext[2],
                max_r=max_r, max_r_pval=pval[imax], max_r_pos=pos[imax],
                min_r=min_r, min_r_pval=pval[imin], min_r_pos=pos[imin], r_range=max_r - min_r)


def get_minmax_stats_df(motif_id_to_profile_df, mt_method='fdr_tsbky', mt_alpha=0.01, thorough_mt=True):
    """utils.py:448-530: one row per motif, with multiple-testing adjusted p-values.
    thorough_mt adjusts over every (motif, position) p-value and looks the extreme / max / min positions up."""
    ids = list(motif_id_to_profile_df)
    out = pd.DataFrame.from_records([dict(motif_id=k, **get_minmax_stats(motif_id_to_profile_df[k])) for k in ids])
    lookup = None
    if thorough_mt and ids:
        first = motif_id_to_profile_df[ids[0]]
        pcol = 'permutation_positional_r_pval' if 'permutation_positional_r_pval' in first.columns else 'positional_r_pval'
        allp = np.concatenate([motif_id_to_profile_df[k][pcol].to_numpy() for k in ids])
        _, padj = multipletests(allp, alpha=mt_alpha, method=mt_method)
        lookup, off = {}, 0
        for k in ids:
            posk = motif_id_to_profile_df[k]['position'].to_numpy()
            lookup[k] = dict(zip(posk.tolist(), padj[off:off + posk.size].tolist()))
            off += posk.size
    for prefix in ('extreme', 'max', 'min'):
        pcol = f'{prefix}_r_pval'
        if thorough_mt:
            padj = np.array([lookup[k][p] for k, p in zip(out['motif_id'], out[f'{prefix}_r_pos'])]) if ids else np.zeros(0)
        else:
            _, padj = multipletests(out[pcol].to_numpy(), alpha=mt_alpha, method=mt_method)
        idx = list(out.columns).index(pcol)
        out.insert(idx + 1, f'{prefix}_r_sig', padj <= mt_alpha)
        out.insert(idx + 1, f'{prefix}_r_padj', padj)
    if ids and not out['integral_r'].isna().any():
        _, padj = multipletests(out['integral_r_pval'].to_numpy(dtype=float), alpha=mt_alpha, method=mt_method)
        idx = list(out.columns).index('integral_r_pval')
        out.insert(idx + 1, 'integral_r_sig', padj <= mt_alpha)
        out.insert(idx + 1, 'integral_r_padj', padj)
    return out


def results_table(filepaths_df, motif_orientation='+', mt_method='fdr_tsbky', mt_alpha=0.01, thorough_mt=True):
    """Tabular part of results_table_orientation_* (clustering.py:355-410 without images): reads the
    positions_df.pkl of every successful task of one orientation (utils.py:310-355 filters) and summarises."""
    import os
    df = filepaths_df[(filepaths_df['orientation'] == motif_orientation) & (filepaths_df['retval'] == 0)]
    profiles = {}
    for motif_id, outdir in zip(df['motif_id'], df['outdir']):
        path = os.path.normpath(f'{outdir}/positions_df.pkl')
        if os.path.exists(path):
            pdf = pd.read_pickle(path)
            if pdf['positional_r'].fillna(0.0).abs().sum() > 0.0:      # utils.py:350-355 drops empty profiles
                profiles[motif_id] = pdf
    if not profiles:
        return pd.DataFrame()
    return get_minmax_stats_df(profiles, mt_method=mt_method, mt_alpha=mt_alpha, thorough_mt=thorough_mt)

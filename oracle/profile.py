"""Oracle restatement of the reference's per-task profiling path (L1).  Test infrastructure only.

Follows mepp/motif_layers.py:12-165 (PWM -> Conv1D weights/bias, padding, ReLU,
dual max), mepp/core.py:59-145 (margin blur), mepp/core.py:182-267 (positional
Pearson / partial correlation), mepp/core.py:269-376 (permutation CI / p-values),
mepp/core.py:380-413 (parametric CI), mepp/core.py:415-470 (counts, density) and
the driver mepp/core.py:490-632.

Canonical float32 summation order of the scan (the reference's TensorFlow order
is unknowable here; one-hot products are exact so only the order of additions
matters -- SURVEY.md section 7 "hard parts"):

    acc = 0f
    for j in 0..m-1:                       # taps, ascending
        base in ACGT:  acc = fl32(acc + W32[base][j])
        base is N   :  acc = fl32(fl32(fl32(fl32(acc + .25*W32[A][j]) + .25*W32[C][j]) + .25*W32[G][j]) + .25*W32[T][j])
    s = fl32(acc + fl32(-thr))

which is what a sequential-k FMA contraction over the inner index (j, channel)
(Eigen's GEBP, the TF CPU Conv path) produces for one-hot / 0.25 inputs.
"""
import warnings
import numpy as np
from scipy import stats as _st

from . import moods


# ----------------------------------------------------------------------------- weights
def motif_weights(pwm, orientation, pseudocount=1e-4, pvalue=1e-4, bg=None,
                  precision=moods.DP_PRECISION):
    """motif_layers.py:12-78.  Returns (W_fwd32, W_rev32 or None, thr64).

    '+'  : filter = log_odds(pwm);            thr from that matrix
    '-'  : filter = log_odds(pwm)[::-1,::-1]; thr from the reversed matrix (:36-48)
    else : two filters (W, W[::-1,::-1]), both with bias -thr(W)   (:66-74)
    The returned "fwd" filter is always the first Conv1D filter.
    """
    if bg is None:
        bg = moods.flat_bg(4)
    W = moods.log_odds(np.asarray(pwm, dtype=np.float64), bg, pseudocount)
    if orientation == '+':
        thr = moods.threshold_from_p(W, bg, pvalue, precision)
        return W.astype(np.float32), None, thr
    if orientation == '-':
        Wr = W[::-1, ::-1]
        thr = moods.threshold_from_p(Wr, bg, pvalue, precision)
        return np.ascontiguousarray(Wr).astype(np.float32), None, thr
    thr = moods.threshold_from_p(W, bg, pvalue, precision)
    return W.astype(np.float32), np.ascontiguousarray(W[::-1, ::-1]).astype(np.float32), thr


# ----------------------------------------------------------------------------- scan
def _conv_valid_f32(codes, W32):
    """Canonical-order float32 'valid' convolution.  codes (N, L) int8 in 0..4."""
    N, L = codes.shape
    m = W32.shape[1]
    nout = L - m + 1
    acc = np.zeros((N, nout), dtype=np.float32)
    q = (np.float32(0.25) * W32).astype(np.float32)  # exact (power of two)
    for j in range(m):
        c = codes[:, j:j + nout]
        is_n = c >= 4
        normal = acc + W32[np.minimum(c, 3), j]
        if is_n.any():
            an = acc + q[0, j]
            an = an + q[1, j]
            an = an + q[2, j]
            an = an + q[3, j]
            acc = np.where(is_n, an, normal)
        else:
            acc = normal
    return acc


def scan(codes, Wf32, Wr32, thr):
    """motif_layers.py:80-165: conv (+ max over the two filters), zero pad, ReLU.

    Returns X0 (N, L) float32, the unsmoothed thresholded score matrix.
    """
    N, L = codes.shape
    m = Wf32.shape[1]
    bias = np.float32(-thr)  # np.zeros(...) - thr stored as float32 weights (:74-76)
    s = _conv_valid_f32(codes, Wf32) + bias
    if Wr32 is not None:
        s = np.maximum(s, _conv_valid_f32(codes, Wr32) + bias)
    pad_left = (m - 1) // 2  # (:99-103)
    X0 = np.zeros((N, L), dtype=np.float32)
    X0[:, pad_left:pad_left + s.shape[1]] = np.maximum(s, np.float32(0))
    return X0


def blur(X0, margin, dtype=np.float64):
    """core.py:59-145: AveragePooling1D(k=1+2*margin, stride 1, 'valid') then zero pad `margin` each side."""
    if margin <= 0:
        return X0.astype(dtype)
    N, L = X0.shape
    k = 1 + 2 * margin
    X = X0.astype(dtype)
    out = np.zeros((N, L), dtype=dtype)
    if L - k + 1 <= 0:
        return out
    acc = X[:, 0:L - k + 1].copy()
    for d in range(1, k):
        acc = acc + X[:, d:L - k + 1 + d]
    out[:, margin:L - margin] = acc / dtype(k)
    return out


# ----------------------------------------------------------------------------- correlation
def _r_ab(sa, sb, saa, sbb, sab, n):
    """core.py:182-193."""
    with np.errstate(invalid='ignore', divide='ignore'):
        return ((n * sab) - (sa * sb)) / np.sqrt(((n * saa) - (sa * sa)) * ((n * sbb) - (sb * sb)))


def positional_correlation(X, Y, z=None):
    """core.py:195-267 in float64.  X (N, L), Y (N, C), z (N,) or None -> r (L, C) float64."""
    X = X.astype(np.float64)
    Y = Y.astype(np.float64)
    n = X.shape[0]
    sx = X.sum(0)[:, None]
    sxx = (X * X).sum(0)[:, None]
    sy = Y.sum(0)[None, :]
    syy = (Y * Y).sum(0)[None, :]
    sxy = X.T @ Y
    r_xy = _r_ab(sx, sy, sxx, syy, sxy, n)
    if z is None:
        return r_xy
    z = z.astype(np.float64)
    sz = z.sum()
    szz = (z * z).sum()
    sxz = (X * z[:, None]).sum(0)[:, None]
    syz = (Y * z[:, None]).sum(0)[None, :]
    r_xz = _r_ab(sx, sz, sxx, szz, sxz, n)
    r_yz = _r_ab(sy, sz, syy, szz, syz, n)
    with np.errstate(invalid='ignore', divide='ignore'):
        return (r_xy - r_xz * r_yz) / np.sqrt((1 - r_xz * r_xz) * (1 - r_yz * r_yz))


# ----------------------------------------------------------------------------- statistics
def positional_r_columns(r, center=None, confidence_interval_pct=95, permutation_mode='semi'):
    """core.py:269-376 on r (L, C) float32 (already nan_to_num'ed).  Returns an ordered dict of columns."""
    r_ = r if r.ndim > 1 else r[:, None]
    L = r_.shape[0]
    has_perm = r_.shape[1] > 1
    cols = {}
    cols['positional_r'] = r_[:, 0].copy()
    if center is None:
        center = L // 2
    cols['position'] = np.arange(L) - center
    if not has_perm:
        return cols
    alpha = 1.0 - (confidence_interval_pct / 100.0)
    q = (100.0 * alpha / 2, 100.0 - 100.0 * alpha / 2)
    perm = r_[:, 1:]
    pct = np.percentile(perm, q, axis=-1)
    pvals = np.mean((np.abs(perm) >= np.abs(r_[:, 0:1])).astype(float), axis=-1)
    flat = perm.flatten()
    semi_pct = np.percentile(flat, q)
    # reference builds an (L, L*P) comparison matrix (core.py:318-324); identical counts via sort
    sabs = np.sort(np.abs(flat))
    t = np.abs(r_[:, 0])
    semi_p = (sabs.size - np.searchsorted(sabs, t, side='left')) / float(sabs.size)
    cols['permutation_full_positional_r_lower'] = pct[0]
    cols['permutation_full_positional_r_upper'] = pct[1]
    cols['permutation_full_positional_r_pval'] = pvals
    cols['permutation_semi_positional_r_lower'] = np.full(L, semi_pct[0])
    cols['permutation_semi_positional_r_upper'] = np.full(L, semi_pct[1])
    cols['permutation_semi_positional_r_pval'] = semi_p
    for suf in ('positional_r_lower', 'positional_r_upper', 'positional_r_pval'):
        cols[f'permutation_{suf}'] = cols[f'permutation_{permutation_mode}_{suf}']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        trapz = getattr(np, 'trapz', None) or np.trapezoid
        integral = trapz(np.abs(r_), axis=0)
    ipct = np.percentile(integral[1:].flatten(), q)
    ipval = np.mean((integral[1:].flatten() >= integral[0]).astype(float), axis=-1)
    cols['integral_r'] = np.full(L, integral[0])
    cols['permutation_integral_r_lower'] = np.full(L, ipct[0])
    cols['permutation_integral_r_upper'] = np.full(L, ipct[1])
    cols['permutation_integral_r_pval'] = np.full(L, ipval)
    return cols


def parametric_ci(r32, num_samples, confidence_interval_pct=95):
    """core.py:380-413.  r is float32; the reference's numpy (<1.24, value-based casting)
    keeps float32 arrays float32 when combined with float64 scalars, restated here."""
    f = np.float32
    r = r32.astype(f)
    n = num_samples
    alpha = 1.0 - (confidence_interval_pct / 100.0)
    z_margin = f(_st.norm.ppf(1.0 - alpha / 2.0) * np.sqrt(1.0 / (n - 3)))
    eps = np.finfo(f).eps
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        z_r = f(0.5) * np.log((f(1) + r + eps) / (f(1) - r + eps))
        z_to_r = lambda z: (np.exp(f(2) * z + eps) - f(1)) / (np.exp(f(2) * z + eps) + f(1))
        lower = z_to_r(z_r - z_margin)
        upper = z_to_r(z_r + z_margin)
        t = r / np.sqrt((f(1) - (r * r)) / f(n - 2))
        pval = _st.t.sf(np.abs(t), n - 1) * 2
    return lower, upper, pval


def counts_by_position(X0):
    """core.py:415-424."""
    return (X0 > 0.0).sum(axis=0).astype(np.float64)


def smooth_counts(count, window):
    """core.py:438-454: centred rolling mean, then ffill / bfill of the NaN edges."""
    L = count.shape[0]
    out = np.full(L, np.nan)
    h = window // 2  # window is odd (1 + 2*margin)
    if L >= window:
        c = np.concatenate([[0.0], np.cumsum(count)])
        out[h:L - h] = (c[window:] - c[:L - window + 1]) / window
        out[:h] = out[h]
        out[L - h:] = out[L - h - 1]
    return out


def density_by_rank(X0):
    """core.py:456-463."""
    return (X0 > 0.0).sum(axis=1).astype(np.float64)


# ----------------------------------------------------------------------------- driver
POSITIONS_COLUMNS = [
    'positional_r', 'position',
    'permutation_full_positional_r_lower', 'permutation_full_positional_r_upper', 'permutation_full_positional_r_pval',
    'permutation_semi_positional_r_lower', 'permutation_semi_positional_r_upper', 'permutation_semi_positional_r_pval',
    'permutation_positional_r_lower', 'permutation_positional_r_upper', 'permutation_positional_r_pval',
    'integral_r', 'permutation_integral_r_lower', 'permutation_integral_r_upper', 'permutation_integral_r_pval',
    'positional_r_lower', 'positional_r_upper', 'positional_r_pval',
    'count', 'smoothed_count',
]


def profile_task(codes, Y32, z32, pwm, orientation='+', margin=0, confidence_interval_pct=95,
                 center=None, precision=moods.DP_PRECISION, blur_dtype=np.float64):
    """core.py:490-632 for one (motif, orientation).

    codes (N, L) int8 0..4, Y32 (N, C) float32 (column 0 true scores), z32 (N,) float32 or None.
    Returns dict(X0, r, positions (ordered dict of columns), ranks (dict), thr).
    """
    N, L = codes.shape
    if center is None:
        center = L // 2
    if center < 0:
        center = 0
    if center > L:
        center = L - 1
    Wf, Wr, thr = motif_weights(pwm, orientation, precision=precision)
    X0 = scan(codes, Wf, Wr, thr)
    X = blur(X0, margin, dtype=blur_dtype)
    r64 = positional_correlation(X, Y32, z32)
    r32 = np.nan_to_num(r64.astype(np.float32), 0.0)  # core.py:602
    cols = positional_r_columns(r32, center=center, confidence_interval_pct=confidence_interval_pct)
    lo, up, pv = parametric_ci(cols['positional_r'], N, confidence_interval_pct)
    cols['positional_r_lower'] = lo
    cols['positional_r_upper'] = up
    cols['positional_r_pval'] = pv
    count = counts_by_position(X0)
    cols['count'] = count
    cols['smoothed_count'] = smooth_counts(count, 1 + 2 * margin) if margin > 0 else count
    ranks = dict(rank=np.arange(N), score=Y32[:, 0].copy(), density=density_by_rank(X0))
    return dict(X0=X0, X=X, r64=r64, r=r32, positions=cols, ranks=ranks, thr=thr, Wf=Wf, Wr=Wr)

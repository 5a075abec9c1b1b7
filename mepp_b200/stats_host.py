"""Host-side finalisation of the per-task statistics (L-length vectors only).

Turns the order statistics / counts produced by the K3 CUDA kernels into the columns of
the reference's positions_df, following mepp/core.py:269-413 and :438-454:
numpy's "linear" percentile interpolation, permutation p-values = count / n, the
parametric Fisher-z confidence interval and t-test, and the rolling-mean count smoothing.
Everything here is O(T * L) numpy on vectors the GPU already reduced.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy import special as _sp
from scipy import stats as _st

_POOL = None


def _threads():
    from .parallel import host_threads
    return host_threads()


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=_threads())
    return _POOL


def _t_sf2(abs_t, df):
    """2 * scipy.stats.t.sf(|t|, df) (core.py:410), chunked over host threads for large inputs
    (scipy.special.stdtr is the function t.sf evaluates; it releases the GIL)."""
    x = -np.asarray(abs_t, dtype=np.float64)
    if x.size < 32768:
        return _sp.stdtr(df, x) * 2
    flat = x.reshape(-1)
    n = _threads()
    parts = list(_pool().map(lambda c: _sp.stdtr(df, c), np.array_split(flat, n)))
    return (np.concatenate(parts) * 2).reshape(x.shape)

POSITIONS_COLUMNS = [
    'positional_r', 'position',
    'permutation_full_positional_r_lower', 'permutation_full_positional_r_upper', 'permutation_full_positional_r_pval',
    'permutation_semi_positional_r_lower', 'permutation_semi_positional_r_upper', 'permutation_semi_positional_r_pval',
    'permutation_positional_r_lower', 'permutation_positional_r_upper', 'permutation_positional_r_pval',
    'integral_r', 'permutation_integral_r_lower', 'permutation_integral_r_upper', 'permutation_integral_r_pval',
    'positional_r_lower', 'positional_r_upper', 'positional_r_pval',
    'count', 'smoothed_count',
]
POSITIONS_COLUMNS_NO_PERM = ['positional_r', 'position', 'positional_r_lower', 'positional_r_upper',
                             'positional_r_pval', 'count', 'smoothed_count']


def percentile_q(confidence_interval_pct):
    """core.py:297-300."""
    alpha = 1.0 - (confidence_interval_pct / 100.0)
    return (100.0 * alpha / 2, 100.0 - 100.0 * alpha / 2)


def virtual_index(n, q_pct):
    """numpy 'linear' method on a float32 array: h = (n-1) * (q / float32(100)); returns (floor, gamma)."""
    quant = np.true_divide(np.float64(q_pct), np.float32(100))
    h = (n - 1) * quant
    k = int(np.floor(h))
    k = min(max(k, 0), n - 1)
    return k, float(h - k)


def lerp_f32(a, b, gamma):
    """numpy's _lerp for float32 neighbours a, b and float64 gamma (difference taken in float32)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    diff = np.subtract(b, a)  # float32, like numpy
    out = a.astype(np.float64) + diff.astype(np.float64) * gamma
    if gamma >= 0.5:
        out = b.astype(np.float64) - diff.astype(np.float64) * (1 - gamma)
    return out


def parametric_ci(r32, num_samples, confidence_interval_pct=95):
    """core.py:380-413 with the float32 arithmetic of the reference's numpy (value-based casting)."""
    f = np.float32
    r = np.asarray(r32, dtype=f)
    n = num_samples
    alpha = 1.0 - (confidence_interval_pct / 100.0)
    z_margin = f(_st.norm.ppf(1.0 - alpha / 2.0) * np.sqrt(1.0 / (n - 3)))
    eps = np.finfo(f).eps
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        z_r = f(0.5) * np.log((f(1) + r + eps) / (f(1) - r + eps))
        lower = (np.exp(f(2) * (z_r - z_margin) + eps) - f(1)) / (np.exp(f(2) * (z_r - z_margin) + eps) + f(1))
        upper = (np.exp(f(2) * (z_r + z_margin) + eps) - f(1)) / (np.exp(f(2) * (z_r + z_margin) + eps) + f(1))
        t = r / np.sqrt((f(1) - (r * r)) / f(n - 2))
        pval = _t_sf2(np.abs(t), n - 1)
    return lower, upper, pval


def smooth_counts(count, window):
    """core.py:438-454 along the last axis: centred rolling mean, NaN edges ffill/bfill-ed."""
    count = np.asarray(count, dtype=np.float64)
    L = count.shape[-1]
    out = np.full(count.shape, np.nan)
    h = window // 2
    if L >= window:
        c = np.concatenate([np.zeros(count.shape[:-1] + (1,)), np.cumsum(count, axis=-1)], axis=-1)
        out[..., h:L - h] = (c[..., window:] - c[..., :L - window + 1]) / window
        out[..., :h] = out[..., h:h + 1]
        out[..., L - h:] = out[..., L - h - 1:L - h]
    return out


PER_POSITION = {  # column -> dtype of the per-(task, position) arrays the host materialises
    'positional_r': np.float32,
    'permutation_full_positional_r_lower': np.float64, 'permutation_full_positional_r_upper': np.float64,
    'permutation_full_positional_r_pval': np.float64, 'permutation_semi_positional_r_pval': np.float64,
    'positional_r_lower': np.float32, 'positional_r_upper': np.float32, 'positional_r_pval': np.float64,
    'count': np.float64, 'smoothed_count': np.float64,
}
PER_TASK = ['permutation_semi_positional_r_lower', 'permutation_semi_positional_r_upper', 'integral_r',
            'permutation_integral_r_lower', 'permutation_integral_r_upper', 'permutation_integral_r_pval']
_PERM_ONLY = {'permutation_full_positional_r_lower', 'permutation_full_positional_r_upper',
              'permutation_full_positional_r_pval', 'permutation_semi_positional_r_pval'}


def alloc_positions(T, L, P):
    """Output store for T tasks: per-position arrays + per-task scalars (broadcast columns are never
    materialised as (T, L) arrays)."""
    pos = {k: np.empty((T, L), dtype=dt) for k, dt in PER_POSITION.items() if P > 0 or k not in _PERM_ONLY}
    task = {k: np.empty(T, dtype=np.float32 if k == 'integral_r' else np.float64) for k in PER_TASK} if P > 0 else {}
    return dict(pos=pos, task=task, T=T, L=L, P=P)


def positions_view(store, center=None, permutation_mode='semi'):
    """{column: (T, L) array or zero-copy broadcast view} in the reference's column order (core.py:286-374,
    406-411, 617-618)."""
    T, L, P = store['T'], store['L'], store['P']
    if center is None:
        center = L // 2
    cols = dict(store['pos'])
    cols['position'] = np.broadcast_to(np.arange(L) - center, (T, L))
    if P > 0:
        for k in PER_TASK:
            cols[k] = np.broadcast_to(store['task'][k][:, None], (T, L))
        for suf in ('positional_r_lower', 'positional_r_upper', 'positional_r_pval'):
            cols[f'permutation_{suf}'] = cols[f'permutation_{permutation_mode}_{suf}']
    order = POSITIONS_COLUMNS if P > 0 else POSITIONS_COLUMNS_NO_PERM
    return {k: cols[k] for k in order}


def finalize_into(store, t0, r0, full_os, full_cnt, semi_os, semi_cnt, integral, count, n_seq, margin,
                  confidence_interval_pct=95):
    """Finalise one group of tasks (rows t0 .. t0+T of the store) from the GPU's order statistics / counts.

    r0 (T, L) f32; full_os (T, L, 4) f32; full_cnt (T, L); semi_os (T, 4) f32; semi_cnt (T, L);
    integral (T, 1+P) f64; count (T, L).  The inputs may be pinned staging buffers: everything is copied out.
    """
    T, L = r0.shape
    P = store['P']
    sl = slice(t0, t0 + T)
    pos, task = store['pos'], store['task']
    pos['positional_r'][sl] = r0
    if P > 0:
        q = percentile_q(confidence_interval_pct)
        (_, g_lo), (_, g_hi) = virtual_index(P, q[0]), virtual_index(P, q[1])
        pos['permutation_full_positional_r_lower'][sl] = lerp_f32(full_os[..., 0], full_os[..., 1], g_lo)
        pos['permutation_full_positional_r_upper'][sl] = lerp_f32(full_os[..., 2], full_os[..., 3], g_hi)
        np.divide(full_cnt, float(P), out=pos['permutation_full_positional_r_pval'][sl])
        (_, gs_lo), (_, gs_hi) = virtual_index(L * P, q[0]), virtual_index(L * P, q[1])
        task['permutation_semi_positional_r_lower'][sl] = lerp_f32(semi_os[:, 0], semi_os[:, 1], gs_lo)
        task['permutation_semi_positional_r_upper'][sl] = lerp_f32(semi_os[:, 2], semi_os[:, 3], gs_hi)
        np.divide(semi_cnt, float(L * P), out=pos['permutation_semi_positional_r_pval'][sl])
        # np.trapz on the float32 r array yields float32 (core.py:352); percentiles over the P permuted integrals
        integ32 = integral.astype(np.float32)
        ipct = np.percentile(integ32[:, 1:], q, axis=-1)
        task['integral_r'][sl] = integ32[:, 0]
        task['permutation_integral_r_lower'][sl] = ipct[0]
        task['permutation_integral_r_upper'][sl] = ipct[1]
        task['permutation_integral_r_pval'][sl] = np.mean((integ32[:, 1:] >= integ32[:, 0:1]).astype(float), axis=-1)
    lo, up, pv = parametric_ci(pos['positional_r'][sl], n_seq, confidence_interval_pct)
    pos['positional_r_lower'][sl] = lo
    pos['positional_r_upper'][sl] = up
    pos['positional_r_pval'][sl] = pv
    pos['count'][sl] = count
    pos['smoothed_count'][sl] = smooth_counts(pos['count'][sl], 1 + 2 * margin) if margin > 0 else count


def finalize_positions(r0, full_os, full_cnt, semi_os, semi_cnt, integral, count, n_seq, P, margin,
                       confidence_interval_pct=95, center=None, permutation_mode='semi'):
    """One-shot form: {column: (T, L) array} in reference order for T tasks (see finalize_into)."""
    T, L = r0.shape
    store = alloc_positions(T, L, P)
    finalize_into(store, 0, r0, full_os, full_cnt, semi_os, semi_cnt, integral, count, n_seq, margin,
                  confidence_interval_pct)
    return positions_view(store, center, permutation_mode)


FINALIZED_F64 = ['permutation_full_positional_r_lower', 'permutation_full_positional_r_upper',
                 'permutation_full_positional_r_pval', 'permutation_semi_positional_r_pval', 'positional_r_pval',
                 'count', 'smoothed_count']
FINALIZED_F32 = ['positional_r', 'positional_r_lower', 'positional_r_upper']
FINALIZED_TASK_F64 = ['permutation_semi_positional_r_lower', 'permutation_semi_positional_r_upper',
                      'permutation_integral_r_lower', 'permutation_integral_r_upper', 'permutation_integral_r_pval']


def z_margin_f32(num_samples, confidence_interval_pct=95):
    """float32(norm.ppf(1 - alpha / 2) * sqrt(1 / (n - 3))) of core.py:393-397."""
    alpha = 1.0 - (confidence_interval_pct / 100.0)
    return np.float32(_st.norm.ppf(1.0 - alpha / 2.0) * np.sqrt(1.0 / (num_samples - 3)))


def store_finalized(store, t0, buf, T, L, a=0, b=None):
    """Copy tasks a .. b of the output block of mepp_finalize_positions (include/mepp_b200.h, K4) of one task group
    of T tasks into rows t0+a .. t0+b of the store.  buf: uint8 array (may be a pinned staging buffer)."""
    P = store['P']
    TL = T * L
    b = T if b is None else b
    sl = slice(t0 + a, t0 + b)
    pos, task = store['pos'], store['task']
    d64 = buf[:TL * 56].view(np.float64).reshape(7, T, L)
    f32 = buf[TL * 56:TL * 68].view(np.float32).reshape(3, T, L)
    for i, k in enumerate(FINALIZED_F64):
        if k in pos:
            pos[k][sl] = d64[i, a:b]
    for i, k in enumerate(FINALIZED_F32):
        pos[k][sl] = f32[i, a:b]
    if P > 0:
        off = (TL * 68 + 7) & ~7
        t64 = buf[off:off + T * 40].view(np.float64).reshape(5, T)
        for i, k in enumerate(FINALIZED_TASK_F64):
            task[k][sl] = t64[i, a:b]
        task['integral_r'][sl] = buf[off + T * 40:off + T * 44].view(np.float32)[a:b]

"""CPU baseline: the reference's algorithm with the reference's structure, on host cores.

TEST INFRASTRUCTURE / BASELINE ONLY (bench.py's cpu_baseline leg and `--impl reference`).
The reference itself (TensorFlow + MOODS) cannot run in this image, so this port is what gets
timed: float32 scan + blur (C restatement, threaded over sequences), then -- batch by batch of
`batch_size` sequences as mepp/core.py:222-253 does -- the nine raw-moment sums in float32, with
the (B, L, 1+P) outer-product reduce_sum of core.py:250 replaced by the mathematically identical
X_b^T @ Y_b float32 matmul (BLAS, all cores; the literal outer product would need 4 GB per
batch), then the numpy statistics of core.py:269-413 (semi p-values by sort + searchsorted).
"""
import os
import time

import numpy as np

from . import cref, profile


def cpu_profile_task(codes, Y32, z32, pwm, orientation, margin, confidence_interval_pct=95, batch_size=1000,
                     threads=None):
    threads = threads or os.cpu_count() or 1
    N, L = codes.shape
    Wf, Wr, thr = profile.motif_weights(pwm, orientation)
    X0 = cref.scan(codes, Wf, Wr, thr, threads=threads)
    X = cref.blur(X0, margin, threads=threads) if margin > 0 else X0
    f = np.float32
    C = Y32.shape[1]
    sx = np.zeros((L, 1), f); sxx = np.zeros((L, 1), f); sxz = np.zeros((L, 1), f)
    sy = np.zeros((1, C), f); syy = np.zeros((1, C), f); syz = np.zeros((1, C), f)
    sz = f(0); szz = f(0)
    sxy = np.zeros((L, C), f)
    for b0 in range(0, N, batch_size):
        xb, yb, zb = X[b0:b0 + batch_size], Y32[b0:b0 + batch_size], z32[b0:b0 + batch_size]
        sx += xb.sum(0, dtype=f)[:, None]; sxx += (xb * xb).sum(0, dtype=f)[:, None]
        sy += yb.sum(0, dtype=f)[None, :]; syy += (yb * yb).sum(0, dtype=f)[None, :]
        sz += zb.sum(dtype=f); szz += (zb * zb).sum(dtype=f)
        sxy += xb.T @ yb
        sxz += (xb.T @ zb)[:, None]
        syz += (zb @ yb)[None, :]
    n = f(N)
    with np.errstate(invalid='ignore', divide='ignore'):
        r_ab = lambda a, b, aa, bb, ab: ((n * ab) - (a * b)) / np.sqrt(((n * aa) - (a * a)) * ((n * bb) - (b * b)))
        r_xy, r_xz, r_yz = r_ab(sx, sy, sxx, syy, sxy), r_ab(sx, sz, sxx, szz, sxz), r_ab(sy, sz, syy, szz, syz)
        r = (r_xy - r_xz * r_yz) / np.sqrt((1 - r_xz * r_xz) * (1 - r_yz * r_yz))
    r = np.nan_to_num(r.astype(f), 0.0)
    cols = profile.positional_r_columns(r, confidence_interval_pct=confidence_interval_pct)
    lo, up, pv = profile.parametric_ci(cols['positional_r'], N, confidence_interval_pct)
    cols.update(positional_r_lower=lo, positional_r_upper=up, positional_r_pval=pv)
    count = profile.counts_by_position(X0)
    cols['count'] = count
    cols['smoothed_count'] = profile.smooth_counts(count, 1 + 2 * margin) if margin > 0 else count
    return cols, profile.density_by_rank(X0)


def time_sample(codes, Y32, z32, tasks, margin, threads=None, budget_s=25.0):
    """Times whole tasks until `budget_s` is spent (at least one).  Returns (seconds, n_tasks_done)."""
    done = 0
    t0 = time.perf_counter()
    for pwm, orient in tasks:
        cpu_profile_task(codes, Y32, z32, pwm, orient, margin, threads=threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return time.perf_counter() - t0, done

"""Restatement of the two MOODS-python functions on the hot path.

Reference call sites: mepp/motif_layers.py:23 (flat_bg), :37-41 (log_odds),
:44-48 (threshold_from_p, 3-argument form).  MOODS-python >= 1.9.4
(setup.py:19) is a SWIG-wrapped C++ library that is NOT under /root/reference;
what follows restates its published algorithm (moods_tools.cpp) from memory --
see SURVEY.md section 8(c).  Test infrastructure only.
"""
import numpy as np

# MOODS multiplies log-odds by this factor before rounding to integers for the
# p-value DP.  SURVEY.md section 8(c) fixes 1000.0 (PVAL_DP_MULTIPLIER); some
# MOODS 1.9.4.x builds are believed to use DEFAULT_DP_PRECISION = 2000.0.  It
# is a parameter everywhere so either can be selected; 1000.0 is the contract.
DP_PRECISION = 2000.0


def flat_bg(alphabet_size=4):
    """MOODS.tools.flat_bg (motif_layers.py:23)."""
    return [1.0 / alphabet_size] * alphabet_size


def log_odds(mat, bg, ps):
    """MOODS.tools.log_odds(mat, bg, ps): natural-log odds with pseudocount.

    col = sum_j (mat[j][i] + ps*bg[j]);  W[j][i] = ln((mat[j][i]+ps*bg[j])/col) - ln(bg[j])
    """
    mat = np.asarray(mat, dtype=np.float64)
    a, n = mat.shape
    out = np.zeros((a, n), dtype=np.float64)
    for i in range(n):
        col = 0.0
        for j in range(a):
            col += mat[j, i] + ps * bg[j]
        for j in range(a):
            out[j, i] = np.log((mat[j, i] + ps * bg[j]) / col) - np.log(bg[j])
    return out


def _int_matrix(pssm, precision):
    pssm = np.asarray(pssm, dtype=np.float64)
    v = precision * pssm
    # C cast (long) truncates toward zero
    return np.where(pssm > 0.0, np.trunc(v + 0.5), np.trunc(v - 0.5)).astype(np.int64)


def threshold_from_p(pssm, bg, p, precision=DP_PRECISION):
    """MOODS.tools.threshold_from_p(pssm, bg, p): score threshold for p-value p.

    Integer DP over the score distribution on a 1/precision grid; returns the
    smallest grid score whose upper-tail probability is <= p.
    """
    mat = _int_matrix(pssm, precision)
    a, n = mat.shape
    maxT = int(mat.max(axis=0).sum())
    minV = int(mat.min())
    R = maxT - n * minV
    table0 = np.zeros(R + 1, dtype=np.float64)
    for j in range(a):
        table0[mat[j, 0] - minV] += bg[j]
    for i in range(1, n):
        table1 = np.zeros(R + 1, dtype=np.float64)
        for j in range(a):
            s = int(mat[j, i] - minV)
            # table1[r + s] += bg[j] * table0[r]  (r ascending, j outer)
            table1[s:] += bg[j] * table0[:R + 1 - s]
        table0 = table1
    # cumulative from the top, sequential
    tail = np.cumsum(table0[::-1])
    idx = np.nonzero(tail > p)[0]
    if idx.size == 0:
        return (n * minV) / precision
    r = R - int(idx[0])
    return (r + n * minV + 1) / precision


def threshold_bruteforce(pssm, bg, p, precision=DP_PRECISION):
    """Independent check of the DP: enumerate all a^n integer scores (n small)."""
    mat = _int_matrix(pssm, precision)
    a, n = mat.shape
    scores = np.zeros(1, dtype=np.int64)
    probs = np.ones(1, dtype=np.float64)
    for i in range(n):
        scores = (scores[:, None] + mat[None, :, i]).ravel()
        probs = (probs[:, None] * np.asarray(bg)[None, :]).ravel()
    order = np.argsort(-scores, kind="stable")
    s, pr = scores[order], probs[order]
    # group equal scores
    uniq, start = np.unique(-s, return_index=True)
    uniq = -uniq
    tail = np.add.reduceat(pr, start).cumsum()
    k = np.nonzero(tail > p)[0]
    if k.size == 0:
        return (n * int(mat.min())) / precision
    return (int(uniq[k[0]]) + 1) / precision

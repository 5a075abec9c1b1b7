"""Prototype (numpy, CPU) of the match-driven form of the scan's phase B that DESIGN.md section 8b names as the next
structural step: the non-zeros of the blurred score matrix built from the sorted match list of one task instead of
from per-chunk staging rows.  Checks itself against the dense float32 blur (sequential sum over the window in
position order, IEEE division -- the order scan.cu's store_chunk uses): identical bits, because adding the zeros
of the dense window changes nothing.

    x_bar[n, p] = fl32( fl32( sum_{d = -margin .. margin} x0[n, p + d] ) / k ),  margin <= p <= L - 1 - margin, else 0

Run: python tools/proto_match_blur.py
"""
import numpy as np


def dense_blur_f32(X0, margin):
    """Reference: core.py:59-145 in float32, sequential sum in position order."""
    N, L = X0.shape
    k = 2 * margin + 1
    out = np.zeros((N, L), dtype=np.float32)
    if L - k + 1 <= 0:
        return out
    acc = np.zeros((N, L - k + 1), dtype=np.float32)
    for d in range(k):
        acc = (acc + X0[:, d:d + L - k + 1]).astype(np.float32)
    nz = acc != 0
    acc[nz] = (acc[nz] / np.float32(k)).astype(np.float32)
    out[:, margin:L - margin] = acc
    return out


def entries_from_matches(seq, pos, x0, L, margin):
    """Match list of one task (any order) -> (seq, position, x_bar) records of the non-zeros of the blurred matrix.
    Matches are sorted by (sequence, position); every match touches the rows pos - margin .. pos + margin inside the
    interior [margin, L - 1 - margin]; a row's value is the float32 sum, in position order, of the matches of the same
    sequence inside its window, divided by k."""
    k = 2 * margin + 1
    order = np.lexsort((pos, seq))
    seq, pos, x0 = seq[order], pos[order], x0[order].astype(np.float32)
    out = []
    i, n = 0, len(seq)
    while i < n:
        j = i
        while j < n and seq[j] == seq[i]:
            j += 1
        # one sequence: matches i .. j-1 in increasing position; walk the rows they touch with a sliding window
        lo = hi = i
        p_first = max(margin, int(pos[i]) - margin)
        p_last = min(L - 1 - margin, int(pos[j - 1]) + margin)
        p = p_first
        while p <= p_last:
            while lo < j and pos[lo] < p - margin:
                lo += 1
            if hi < lo:
                hi = lo
            while hi < j and pos[hi] <= p + margin:
                hi += 1
            if lo == hi:
                p = int(pos[lo]) - margin if lo < j else p_last + 1      # gap: jump to the next match's first row
                continue
            acc = np.float32(0.0)
            for q in range(lo, hi):
                acc = np.float32(acc + x0[q])
            if acc != 0:
                out.append((int(seq[i]), p, np.float32(acc / np.float32(k))))
            p += 1
        i = j
    return out


def _check(N=96, L=211, margin=2, rate=3e-3, seed=0):
    rng = np.random.default_rng(seed)
    X0 = np.zeros((N, L), dtype=np.float32)
    hit = rng.random((N, L)) < rate
    hit[3, 50:60] = True                       # a run of overlapping matches
    hit[5, :3] = True                          # matches inside the zeroed edge
    hit[5, L - 2:] = True
    X0[hit] = rng.gamma(2.0, 1.5, size=int(hit.sum())).astype(np.float32) + np.float32(1e-3)
    seq, pos = np.nonzero(X0)
    perm = rng.permutation(len(seq))
    rec = entries_from_matches(seq[perm], pos[perm], X0[seq, pos][perm], L, margin)
    Xb = np.zeros((N, L), dtype=np.float32)
    for s, p, v in rec:
        assert Xb[s, p] == 0
        Xb[s, p] = v
    ref = dense_blur_f32(X0, margin)
    assert np.array_equal(Xb.view(np.uint32), ref.view(np.uint32)), (margin, np.abs(Xb - ref).max())
    return len(rec)


if __name__ == "__main__":
    for mg in (0, 1, 2, 5, 16):
        print(f"margin {mg}: {_check(margin=mg, seed=mg)} non-zeros, bit-identical to the dense float32 blur")

"""Heat-map reduction of the (N, L) motif score matrix straight from the match list.

The reference ships the dense float32 (N, L) matrix to the host for every task
(mepp/core.py:568-573, 400 MB per task at 100k x 1 kb) only so that mepp/plot.py can clip it to
mean + sigma * std of the NON-ZERO scores (plot.py:57-66) and block-max it down to the figure's
pixel grid (plot.py:32-55, skimage.measure.block_reduce with np.max, zero padded).  Both steps only
depend on the matches, so here they run on the engine's match list ({seq, pos, x0} records, a few
hundred KB per task) and the dense matrix never exists.
"""
import numpy as np


def get_aspect_tensor(shape, pixel_width, pixel_height):
    """plot.py:32-40: block size (rows, cols) of the reduction."""
    aspect = np.array([shape[0] / pixel_height, shape[1] / pixel_width])
    aspect = np.clip(aspect, 1.0, np.max(aspect))
    aspect = tuple(np.round(aspect).astype(int))
    return tuple(max(1, int(x)) for x in aspect)


def clip_limit(x0, sigma=1.0):
    """plot.py:57-66: upper clip = mean + sigma * std over the non-zero entries (numpy masked-array
    mean / population std).  x0: the match scores (all > 0).  Returns None for sigma=None."""
    if sigma is None:
        return None
    x0 = np.asarray(x0)
    x0 = x0[x0 != 0]
    if x0.size == 0:
        return 0.0
    m = np.ma.masked_equal(x0, 0.0)
    return float(m.mean() + m.std() * sigma)


def compress_hits_to_pixels(seq, pos, x0, shape, pixel_width, pixel_height, sigma=1.0):
    """Clipped block-max of the score matrix, computed from its non-zero entries.

    seq, pos, x0: the matches of one task; shape = (N, L).  Equals
    compress_matrix_to_pixels(clip_motif_score_matrix(X0, sigma), pixel_width, pixel_height, np.max)
    of the reference (block_reduce pads the last partial blocks with zeros, which never win a max of
    non-negative values)."""
    bh, bw = get_aspect_tensor(shape, pixel_width, pixel_height)
    out = np.zeros((-(-shape[0] // bh), -(-shape[1] // bw)), dtype=np.float32)
    x0 = np.asarray(x0, dtype=np.float32)
    if x0.size:
        lim = clip_limit(x0, sigma)
        vals = x0 if lim is None else np.clip(x0, 0.0, np.float32(lim)).astype(np.float32)
        np.maximum.at(out, (np.asarray(seq) // bh, np.asarray(pos) // bw), vals)
    return out


def compress_dense_to_pixels(X0, pixel_width, pixel_height, sigma=1.0):
    """Dense restatement of the reference's two steps (used by the tests as the checker)."""
    X0 = np.asarray(X0, dtype=np.float32)
    lim = clip_limit(X0[X0 != 0], sigma)
    Xc = X0 if lim is None else np.clip(X0, 0.0, np.float32(lim) if X0.any() else 0.0)
    bh, bw = get_aspect_tensor(X0.shape, pixel_width, pixel_height)
    H, W = -(-X0.shape[0] // bh), -(-X0.shape[1] // bw)
    pad = np.zeros((H * bh, W * bw), dtype=np.float32)
    pad[:X0.shape[0], :X0.shape[1]] = Xc
    return pad.reshape(H, bh, W, bw).max(axis=(1, 3))

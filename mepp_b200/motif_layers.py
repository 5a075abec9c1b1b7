"""Host side of the PWM scan: log-odds weights and the MOODS-derived threshold.

Mirrors mepp/motif_layers.py:12-78 of the reference (motif_matrix_to_conv1d_layer):
instead of a Keras Conv1D layer it returns the scan table consumed by the CUDA scan
kernel -- the same float32 filter weights (one or two filters) and the same bias
-threshold.  The arithmetic runs in the C++ host code of libmepp_b200.so
(mepp_log_odds / mepp_threshold_from_p / mepp_motif_table), as the reference runs
MOODS' C++ on the host.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import os

import numpy as np

from . import _lib

# MOODS' integer grid for the p-value DP (SURVEY.md section 8c fixes 1000.0).
# Grid of the integer DP of MOODS' threshold_from_p: scores are rounded to multiples of 1 / DP_PRECISION.  MOODS >= 1.9.3
# (the reference pins MOODS-python >= 1.9.4, setup.py:19) uses DEFAULT_DP_PRECISION = 2000; older builds used 1000
# (PVAL_DP_MULTIPLIER) -- that is the grid SURVEY.md section 8c restates.  Neither can be checked against MOODS in this
# image; 2000 is the default, MEPP_DP_PRECISION / the dp_precision keyword select the other one, and every task's
# params.yaml records the grid that was used.
DP_PRECISION = float(os.environ.get("MEPP_DP_PRECISION", "2000"))


def flat_bg(alphabet_size=4):
    """MOODS.tools.flat_bg (motif_layers.py:23)."""
    return [1.0 / alphabet_size] * alphabet_size


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def log_odds(motif_matrix, bg=None, pseudocount=0.0001):
    """MOODS.tools.log_odds as called at motif_layers.py:36-41.  Returns (4, m) float64."""
    pwm = np.ascontiguousarray(motif_matrix, dtype=np.float64)
    if pwm.ndim != 2 or pwm.shape[0] != 4:
        raise ValueError("motif_matrix must have shape (4, m)")
    bg = np.asarray(flat_bg(4) if bg is None else bg, dtype=np.float64)
    W = np.empty_like(pwm)
    _lib.check(_lib.lib.mepp_log_odds(_pd(pwm), pwm.shape[1], _pd(bg), float(pseudocount), _pd(W)), "log_odds")
    return W


def threshold_from_p(weights, bg=None, pvalue=0.0001, dp_precision=DP_PRECISION):
    """MOODS.tools.threshold_from_p as called at motif_layers.py:44-48."""
    W = np.ascontiguousarray(weights, dtype=np.float64)
    bg = np.asarray(flat_bg(4) if bg is None else bg, dtype=np.float64)
    thr = C.c_double(0.0)
    _lib.check(_lib.lib.mepp_threshold_from_p(_pd(W), W.shape[1], _pd(bg), float(pvalue), float(dp_precision),
                                              C.byref(thr)), "threshold_from_p")
    return thr.value


def motif_matrix_to_scan_table(motif_matrix, orientation='+', pseudocount=0.0001, pvalue=0.0001, bg=None,
                               honour_flags=False, dp_precision=DP_PRECISION):
    """Scan table for one (motif, orientation) task.

    orientation '+' -> one filter W; '-' -> one filter W[::-1, ::-1] with the threshold of the
    reversed matrix; anything else -> two filters sharing the forward threshold
    (motif_layers.py:36-76, core.py:511-534).  With honour_flags=False the reference's
    behaviour is reproduced: pseudocount / pvalue / bg never reach MOODS (core.py:513-534).
    """
    pwm = np.ascontiguousarray(motif_matrix, dtype=np.float64)
    if pwm.ndim != 2 or pwm.shape[0] != 4:
        raise ValueError("motif_matrix must have shape (4, m)")
    m = pwm.shape[1]
    if not 1 <= m <= _lib.MAX_MOTIF_LEN:
        raise ValueError(f"motif length {m} outside [1, {_lib.MAX_MOTIF_LEN}]")
    och = ord('+') if orientation == '+' else ord('-') if orientation == '-' else ord('d')
    lut = np.zeros((_lib.MAX_MOTIF_LEN, 4, 2), dtype=np.float32)
    bias = C.c_float(0.0)
    nf = C.c_int(0)
    thr = C.c_double(0.0)
    bgp = None if bg is None else _pd(np.asarray(bg, dtype=np.float64))
    qtab = np.zeros(1024, dtype=np.uint32)
    qthr = np.zeros(2, dtype=np.uint32)
    _lib.check(_lib.lib.mepp_motif_table(
        _pd(pwm), m, och, float(pseudocount), float(pvalue), bgp, 1 if honour_flags else 0, float(dp_precision),
        lut.ctypes.data_as(C.POINTER(C.c_float)), C.byref(bias), C.byref(nf), C.byref(thr),
        qtab.ctypes.data, qthr.ctypes.data), "motif_table")
    return dict(lut=lut, bias=np.float32(bias.value), n_filters=nf.value, thr=thr.value, m=m, qtab=qtab, qthr=qthr,
                pvalue=float(pvalue) if honour_flags else 0.0001)


_POOLS = {}


def _pool(n_threads=None):
    """Persistent host thread pool for the threshold DP (ctypes drops the GIL)."""
    from .parallel import host_threads
    n = n_threads or host_threads()
    if n not in _POOLS:
        _POOLS[n] = ThreadPoolExecutor(max_workers=n)
    return _POOLS[n]


def _table_with_threshold(motif_matrix, orientation, thr, pseudocount=0.0001, pvalue=0.0001, bg=None,
                          honour_flags=False, dp_precision=DP_PRECISION):
    """Scan table of a task whose threshold is already known (the '+' and dual orientations of a motif share it)."""
    pwm = np.ascontiguousarray(motif_matrix, dtype=np.float64)
    m = pwm.shape[1]
    och = ord('+') if orientation == '+' else ord('-') if orientation == '-' else ord('d')
    lut = np.zeros((_lib.MAX_MOTIF_LEN, 4, 2), dtype=np.float32)
    bias = C.c_float(0.0)
    nf = C.c_int(0)
    bgp = _pd(np.ascontiguousarray(bg, dtype=np.float64)) if bg is not None else None
    qtab = np.zeros(1024, dtype=np.uint32)
    qthr = np.zeros(2, dtype=np.uint32)
    _lib.check(_lib.lib.mepp_motif_table_with_threshold(
        _pd(pwm), m, och, float(pseudocount), bgp, 1 if honour_flags else 0, float(thr),
        lut.ctypes.data_as(C.POINTER(C.c_float)), C.byref(bias), C.byref(nf), qtab.ctypes.data, qthr.ctypes.data),
        "motif_table_with_threshold")
    return dict(lut=lut, bias=np.float32(bias.value), n_filters=nf.value, thr=float(thr), m=m, qtab=qtab, qthr=qthr,
                pvalue=float(pvalue) if honour_flags else 0.0001)


def scan_tables_async(tasks, n_threads=None, **kw):
    """Submit the scan tables of a list of (motif_matrix, orientation) to the host thread pool (the threshold DP runs
    in C, ctypes drops the GIL) and return a function that waits for them.  One DP per distinct (motif, filter-0
    orientation): the '+' and dual tasks of a motif share the threshold of W (motif_layers.py:36-76)."""
    first = {}
    keys = []
    for pwm, orient in tasks:
        o = orient if orient in ('+', '-') else 'd'
        key = (np.ascontiguousarray(pwm, dtype=np.float64).tobytes(), np.shape(pwm), '-' if o == '-' else 'f')
        keys.append((key, o))
        first.setdefault(key, (pwm, orient))
    pool = _pool(n_threads)
    fut = {key: pool.submit(motif_matrix_to_scan_table, pwm, orient, **kw) for key, (pwm, orient) in first.items()}

    def collect():
        done, out = {}, []
        for (key, o), (pwm, orient) in zip(keys, tasks):
            if (key, o) not in done:
                base = fut[key].result()
                base_o = first[key][1] if first[key][1] in ('+', '-') else 'd'
                done[(key, o)] = base if base_o == o else _table_with_threshold(pwm, orient, base['thr'], **kw)
            out.append(done[(key, o)])
        return out

    return collect


def scan_tables(tasks, n_threads=None, **kw):
    """Scan tables for a list of (motif_matrix, orientation) (see scan_tables_async)."""
    return scan_tables_async(tasks, n_threads, **kw)()


def prepare_tables(tasks, pseudocount=0.0001, pvalue=0.0001, bg=None, honour_flags=False, dp_precision=DP_PRECISION):
    """Everything of the scan tables that does not need the threshold, for a whole task list in one C call, plus the
    inputs of the device threshold DP (include/mepp_b200.h: mepp_prepare_tables).  Returns a dict of numpy arrays;
    `unit_of_task` maps every task to its threshold unit (the '+' and dual orientations of a motif share the
    threshold of W, motif_layers.py:36-76), `unit_rows` are the task rows that represent the units, `twin_of`
    pairs a dual task with the '+' task of the same motif (-1: none)."""
    T = len(tasks)
    pwms = np.zeros((T, 4, _lib.MAX_MOTIF_LEN), dtype=np.float64)
    mlen = np.empty(T, dtype=np.int32)
    ochar = np.empty(T, dtype=np.int32)
    unit_index, unit_of_task, unit_rows = {}, np.empty(T, dtype=np.int32), []
    plus_of, duals = {}, []
    seen = {}                                   # id(matrix object) -> (bytes, m, row): orientations of one motif usually
    for t, (pwm, orient) in enumerate(tasks):   # share the matrix object, which is then packed and hashed once
        hit = seen.get(id(pwm))
        if hit is None:
            a = np.ascontiguousarray(pwm, dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != 4:
                raise ValueError("motif_matrix must have shape (4, m)")
            m = a.shape[1]
            if not 1 <= m <= _lib.MAX_MOTIF_LEN:
                raise ValueError(f"motif length {m} outside [1, {_lib.MAX_MOTIF_LEN}]")
            pwms[t, :, :m] = a
            hit = seen[id(pwm)] = (a.tobytes(), m, t, pwm)      # (the object is kept alive: ids are not reused)
        else:
            pwms[t] = pwms[hit[2]]
        raw, m = hit[0], hit[1]
        mlen[t] = m
        o = orient if orient in ('+', '-') else 'd'
        ochar[t] = ord(o)
        key = (raw, m, '-' if o == '-' else 'f')
        if key not in unit_index:
            unit_index[key] = len(unit_rows)
            unit_rows.append(t)
        unit_of_task[t] = unit_index[key]
        if o == '+':
            plus_of.setdefault(key, []).append(t)
        elif o == 'd':
            duals.append((t, key))
    twin_of = np.full(T, -1, dtype=np.int32)
    for t, key in duals:
        cand = plus_of.get(key)
        if cand:
            twin_of[t] = cand.pop(0)
    out = dict(lut=np.empty((T, _lib.MAX_MOTIF_LEN, 4, 2), dtype=np.float32), nfilt=np.empty(T, dtype=np.int32),
               qtab=np.empty((T, 1024), dtype=np.uint32), qpar=np.empty((T, 4), dtype=np.float64),
               smat=np.empty((T, _lib.MAX_MOTIF_LEN, 4), dtype=np.int32), dp_par=np.empty((T, 8), dtype=np.float64),
               mlen=mlen, unit_of_task=unit_of_task, unit_rows=np.array(unit_rows, dtype=np.int64), twin_of=twin_of,
               pvalue=float(pvalue) if honour_flags else 0.0001)
    bgp = _pd(np.ascontiguousarray(bg, dtype=np.float64)) if bg is not None else None
    _lib.check(_lib.lib.mepp_prepare_tables(pwms.ctypes.data, mlen.ctypes.data, ochar.ctypes.data, T, float(pseudocount),
                                            float(pvalue), bgp, 1 if honour_flags else 0, float(dp_precision),
                                            out['lut'].ctypes.data, out['nfilt'].ctypes.data, out['qtab'].ctypes.data,
                                            out['qpar'].ctypes.data, out['smat'].ctypes.data, out['dp_par'].ctypes.data),
               "prepare_tables")
    return out

"""Writers of the per-task output directories (reference: mepp/single.py:198-302, 342-391).

The reference writes every task directory from its own child process (`python -m mepp.single`), so the
pandas / yaml serialisation of 324 tasks runs on as many cores as there are joblib workers.  Here the compute of
all tasks is one engine call per rank, and the directories are written by a pool of WRITER PROCESSES that runs
beside it: finished task groups are handed over while later groups are still on the GPU.  The numeric tables go
through the native writer (mepp_write_tsv, byte-identical to DataFrame.to_csv: tests/test_host.py), the pickles
through pandas (they must stay pandas pickles for the consumers: utils.py:310-389), the yaml files through
libyaml when PyYAML was built with it.

A worker imports numpy / pandas / yaml and this module only -- never torch or CUDA.
"""
import ctypes as C
import io
import os
import traceback
from os.path import normpath

import numpy as np
import yaml

_KIND = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int64): 2, np.dtype(np.int32): 3,
         np.dtype(np.uint16): 4}
_YAML_DUMPER = getattr(yaml, 'CDumper', yaml.Dumper)


def yaml_dump(obj, path):
    """yaml.dump(obj, f, default_flow_style=False, allow_unicode=True) as single.py:280-302 calls it (libyaml's
    emitter when available: same text, ~10x faster)."""
    with io.open(path, 'w', encoding='utf8') as f:
        yaml.dump(obj, f, Dumper=_YAML_DUMPER, default_flow_style=False, allow_unicode=True)


def write_tsv(path, columns, float_uint16=()):
    """DataFrame(columns).to_csv(path, sep='\\t', index=False) for 1-D numeric columns, through the native writer.
    columns: {name: 1-D array}; a broadcast view (stride 0) is fine.  Names in `float_uint16` are uint16 arrays
    written as float64 values ("3.0").  Other dtypes / names that need quoting take the pandas path."""
    names = list(columns)
    arrs = [np.asarray(columns[k]) for k in names]
    native = all(a.ndim == 1 and a.dtype in _KIND for a in arrs) and \
        all(isinstance(k, str) and k and not any(ch in k for ch in '\t\n\r",') for k in names) and \
        len({a.shape[0] for a in arrs}) == 1
    if not native:
        import pandas as pd
        pd.DataFrame({k: (a.astype(np.float64) if k in float_uint16 else a) for k, a in zip(names, arrs)}).to_csv(
            path, sep='\t', index=False)
        return
    from . import _lib
    n = len(names)
    kinds = (C.c_int32 * n)(*[5 if k in float_uint16 and a.dtype == np.uint16 else _KIND[a.dtype]
                              for k, a in zip(names, arrs)])
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    strides = (C.c_int64 * n)(*[a.strides[0] if a.shape[0] > 0 else 0 for a in arrs])
    _lib.check(_lib.lib.mepp_write_tsv(os.fsencode(path), '\t'.join(names).encode('utf8'), n, kinds, ptrs, strides,
                                       int(arrs[0].shape[0])), "write_tsv")


# ---------------------------------------------------------------------------------------------- task directories
_SHARED = {}


def _init_worker(scores):
    _SHARED['scores'] = scores


def _warm():
    """Start-up work of a writer process, done while the parent still parses its inputs."""
    import pandas  # noqa: F401
    from . import _lib  # noqa: F401
    return 0


def _scores_of(job):
    """The (N,) float32 score column of ranks_df: in the job, or a .npy file every worker reads once."""
    if job.get('scores') is not None:
        return job['scores']
    path = job.get('scores_path')
    if path:
        if path not in _SHARED:
            _SHARED[path] = np.load(path)
        return _SHARED[path]
    return _SHARED['scores']


def write_task_dir(job):
    """Everything single.py:342-391 + :198-302 leave in one task directory.  job: dict(outdir, motif_matrix,
    wrapped_params, params, filepaths, positions {column: (L,) array} or None, density (N,) or None,
    motif_score_matrix or None, save_profile_data, save_motif_score_matrix, error or None).
    Returns (retval, yaml_filepath, log_filepath); never raises (a failure is recorded like a failed child)."""
    outdir = job['outdir']
    yaml_fp, log_fp = normpath(f'{outdir}/wrapped_params.yaml'), normpath(f'{outdir}/wrapped_params.log')
    retval = 0
    log = ''
    try:
        os.makedirs(outdir, exist_ok=True)
        np.save(normpath(f'{outdir}/motif_matrix.npy'), job['motif_matrix'])
        yaml_dump(job['wrapped_params'], yaml_fp)
        if job.get('error'):
            raise RuntimeError(job['error'])
        _write_outputs(job)
    except Exception:  # noqa: BLE001 -- recorded, like the stderr of a failed child process (single.py:392-408)
        log = traceback.format_exc()
        retval = 1
    try:
        with open(log_fp, 'w') as lf:
            lf.write(log)
    except OSError:
        retval = 1
    return retval, yaml_fp, log_fp


def _write_outputs(job):
    import pandas as pd
    fps = dict(job['filepaths'])
    fps.pop('processed_dataset', None)          # save_datasets (TF snapshots of lazy datasets) is not supported
    fps.pop('smoothed_processed_dataset', None)
    if job['save_profile_data']:
        if job['save_motif_score_matrix']:
            np.savez_compressed(fps['motif_score_matrix'], motif_score_matrix=np.asarray(job['motif_score_matrix']))
        else:
            fps.pop('motif_score_matrix', None)
        pos = {k: np.asarray(v) for k, v in job['positions'].items()}
        pd.DataFrame({k: np.array(v) for k, v in pos.items()}).to_pickle(fps['positions_df']['pickle'])
        write_tsv(fps['positions_df']['tsv'], pos)
        dens = np.asarray(job['density'])
        scores = _scores_of(job)
        n = dens.shape[0]
        ranks = dict(rank=np.arange(n), score=scores, density=dens)
        pd.DataFrame(dict(rank=ranks['rank'], score=scores, density=dens.astype(np.float64))).to_pickle(
            fps['ranks_df']['pickle'])
        if dens.dtype == np.uint16:
            write_tsv(fps['ranks_df']['tsv'], ranks, float_uint16=('density',))
        else:
            ranks['density'] = dens.astype(np.float64)
            write_tsv(fps['ranks_df']['tsv'], ranks)
    else:
        for k in ('motif_score_matrix', 'positions_df', 'ranks_df'):
            fps.pop(k, None)
    yaml_dump(job['params'], fps['params'])
    yaml_dump(fps, fps['filepaths'])


class TaskWriter:
    """Pool of writer processes (spawned: they never inherit a CUDA context).  submit() takes the jobs of
    write_task_dir; results() returns their (retval, yaml, log) triples in submission order."""

    def __init__(self, scores=None, n_workers=None, prestart=False):
        """scores: the dataset's (N,) float32 scores, or None when they are not known yet (run_mepp starts the
        writers before it parses its inputs; set_scores() then publishes them through a file).  prestart: spawn
        every worker now instead of with the first jobs."""
        from .parallel import host_threads
        self.scores = None if scores is None else np.ascontiguousarray(scores)
        self.n_workers = host_threads() if n_workers is None else int(n_workers)
        self._pool = None
        self._futures = []
        self._scores_path = None
        if self.n_workers > 1:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor
            self._pool = ProcessPoolExecutor(max_workers=self.n_workers, mp_context=mp.get_context('spawn'),
                                             initializer=_init_worker, initargs=(self.scores,))
            if prestart:
                self._warm = [self._pool.submit(_warm) for _ in range(self.n_workers)]
        else:
            _init_worker(self.scores)

    def set_scores(self, scores):
        scores = np.ascontiguousarray(scores)
        if self._pool is None:
            self.scores = scores
            _init_worker(scores)
            return
        import tempfile
        fd, path = tempfile.mkstemp(prefix='mepp_scores_', suffix='.npy')
        with os.fdopen(fd, 'wb') as f:
            np.save(f, scores)
        self._drop_scores_file()
        self._scores_path = path

    def _drop_scores_file(self):
        if self._scores_path:
            try:
                os.unlink(self._scores_path)
            except OSError:
                pass
            self._scores_path = None

    def submit(self, job):
        if self._scores_path and job.get('scores') is None:
            job = dict(job, scores_path=self._scores_path)
        if self._pool is None:
            self._futures.append(write_task_dir(job))
        else:
            self._futures.append(self._pool.submit(write_task_dir, job))

    def results(self):
        out = []
        for f in self._futures:
            if isinstance(f, tuple):
                out.append(f)
                continue
            try:
                out.append(f.result())
            except Exception:  # noqa: BLE001 -- a dead worker fails its task, not the run
                out.append((1, None, None))
        self._futures = []
        return out

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        self._drop_scores_file()

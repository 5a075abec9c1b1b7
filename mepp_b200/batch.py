"""Task fan-out over motif x orientation (reference: mepp/batch.py).

run_batch keeps mepp/batch.py:28-58's signature and returns the same filepaths_df
(motif_id, orientation, outdir, params, log, retval, attempts; batch.py:174-185) over the same
directory layout (out/<slug(motif_id)>/orientation_{fwd,rev,fwd-rev}/..., batch.py:19-26).
Instead of joblib workers that each spawn `python -m mepp.single` and re-read the dataset from
disk (batch.py:134-148), every task of this process's share runs through ONE engine call
(core.profile_tasks): the sequences are staged in HBM once and tasks are grouped into fused
scan -> GEMM -> statistics launches.  With torch.distributed initialised (one process per GPU) the
task list is sharded over ranks (parallel.shard_tasks); each rank writes its own task directories
and rank 0 gathers the filepaths_df rows.
"""
import os
import re
import shutil
import traceback
import unicodedata
from os.path import normpath

import numpy as np
import pandas as pd

from . import parallel
from .core import profile_tasks
from .io import save_dataset
from .single import _task_filepaths

orientation_to_filepath = {'+': 'fwd', '-': 'rev', '+/-': 'fwd-rev'}


def slugify(text):
    """python-slugify's default behaviour for the ASCII motif ids MEPP uses (the package is not a
    dependency here): NFKD fold to ASCII, lower-case, every run of non-alphanumerics -> '-'."""
    text = unicodedata.normalize('NFKD', str(text)).encode('ascii', 'ignore').decode('ascii')
    text = re.sub(r"[']+", '', text.lower())
    return re.sub(r'[^a-z0-9]+', '-', text).strip('-')


def _task_error(motif_matrix, sequence_length):
    """Why this task cannot run on the engine, or None.  The limits are the kernel's (include/mepp_b200.h)."""
    from ._lib import MAX_MOTIF_LEN
    try:
        m = np.asarray(motif_matrix, dtype=np.float64)
    except (TypeError, ValueError):
        return 'motif matrix is not numeric'
    if m.ndim != 2 or m.shape[0] != 4 or m.shape[1] < 1:
        return f'motif matrix must be 4 x m, got {m.shape}'
    if not np.isfinite(m).all():
        return 'motif matrix holds non-finite values'
    if m.shape[1] > MAX_MOTIF_LEN:
        return f'motif has {m.shape[1]} positions; the scan kernel handles at most {MAX_MOTIF_LEN}'
    if m.shape[1] > sequence_length:
        return f'motif ({m.shape[1]} positions) is longer than the sequences ({sequence_length})'
    return None


motif_id_and_orientation_to_filepath = lambda motif_id, orientation: (
    normpath(f'{slugify(motif_id)}/orientation_{orientation_to_filepath[orientation]}'))


def run_batch(
    dataset, motif_matrix_dict, out_filepath, center=None, motif_orientations=['+', '-', '+/-'], motif_margin=1,
    motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None, confidence_interval_pct=95, motif_score_sigma=1.0,
    motif_score_cmap='gray', rank_smoothing_factor=5, figure_width=10, figure_height=10, save_datasets=False,
    save_profile_data=True, save_motif_score_matrix=False, save_figures=True, figure_formats=['png', 'svg'],
    figure_dpi=300, n_jobs=1, free_dataset=True, keep_dataset=False, progress_wrapper=None, no_gpu=False,
    stop_max_attempt_number=3, wait_random_min=1.0, wait_random_max=2.0, task_writer=None, dp_precision=None
):
    """batch.py:28-190.  `task_writer` (a writer.TaskWriter) is an extension: run_mepp starts the writer processes
    before it parses its inputs and hands them over."""
    if no_gpu:
        from .utils import force_cpu_only
        force_cpu_only()
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    dataset_filepath = normpath(f'{out_filepath}/dataset')
    if rank == 0:
        os.makedirs(dataset_filepath, exist_ok=True)
        # batch.py:60-70 saves the dataset for its child processes and :187-188 deletes it again unless --keepdata;
        # here no child reads it, so the arrays are only written when they are kept (the sidecar file always is)
        save_dataset(dataset, dataset_filepath, delete_existing=True, arrays=bool(keep_dataset))

    tups = [(motif_id, orient,
             normpath(f'{out_filepath}/' + motif_id_and_orientation_to_filepath(motif_id, orient)))
            for motif_id in motif_matrix_dict for orient in motif_orientations]      # batch.py:73-90 order
    tasks = [(motif_matrix_dict[mid], o) for mid, o, _ in tups]
    my_ids = parallel.shard_task_units(tasks, world)[rank]      # orientation pairs of a motif stay on one rank

    # A task that cannot run fails ALONE, like a failed child process of the reference (single.py:392-406): it gets
    # retval 1 and its reason in wrapped_params.log, the other tasks of the rank run.
    L, N = dataset.sequence_length, dataset.num_sequences
    errors = {i: _task_error(tasks[i][0], L) for i in my_ids}
    good = [i for i in my_ids if errors[i] is None]
    attempts = {i: 1 for i in my_ids}

    from .motif_layers import DP_PRECISION
    dp_grid = DP_PRECISION if dp_precision is None else dp_precision     # recorded in every params.yaml
    tw = task_writer
    own_writer = tw is None
    if own_writer:
        from .writer import TaskWriter
        small = len(my_ids) * N < 2_000_000        # (a few small tables: not worth spawning writer processes)
        tw = TaskWriter(dataset.scores, n_workers=0 if small else None)
    submitted = []

    def submit(i, positions=None, density=None, msm=None, error=None):
        motif_id, orient, outdir = tups[i]
        motif_matrix_filepath = normpath(f'{outdir}/motif_matrix.npy')
        params = dict(
            dataset_filepath=dataset_filepath, motif_matrix_filepath=motif_matrix_filepath, motif_id=motif_id,
            output_filepath=outdir, center=center, motif_orientation=orient, motif_margin=motif_margin,
            motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=bg,
            confidence_interval_pct=confidence_interval_pct, motif_score_sigma=motif_score_sigma,
            motif_score_cmap=motif_score_cmap, rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width,
            figure_height=figure_height, save_datasets=save_datasets, save_profile_data=save_profile_data,
            save_motif_score_matrix=save_motif_score_matrix, save_figures=True, figure_formats=list(figure_formats),
            figure_dpi=figure_dpi, no_gpu=no_gpu, dp_precision=float(dp_grid))
        fps = _task_filepaths(outdir, dataset_filepath, motif_matrix_filepath, figure_formats)
        fps.pop('mepp_plot', None)                 # figures: plotting layer, out of scope
        tw.submit(dict(outdir=outdir, motif_matrix=np.asarray(motif_matrix_dict[motif_id]), wrapped_params=params,
                       params={k: v for k, v in params.items() if k not in ('center', 'no_gpu')}, filepaths=fps,
                       positions=positions, density=density, motif_score_matrix=msm,
                       save_profile_data=save_profile_data, save_motif_score_matrix=save_motif_score_matrix,
                       error=error))
        submitted.append(i)

    def run(ids):
        """One engine call over `ids`; finished task groups go to the writers while later groups compute."""
        sel = [tasks[i] for i in ids]

        def on_group(t0, t1, positions, density):
            if save_motif_score_matrix:
                return                              # (the match list is complete only after the last group)
            for j in range(t0, t1):
                submit(ids[j], {k: v[j] for k, v in positions.items()}, density[j])

        batch = profile_tasks(dataset, sel, motif_margin=motif_margin, confidence_interval_pct=confidence_interval_pct,
                              center=center, return_hits=save_motif_score_matrix, on_group=on_group,
                              dp_precision=dp_precision)
        if save_motif_score_matrix:
            for j, i in enumerate(ids):
                submit(i, {k: v[j] for k, v in batch.positions.items()}, batch.density[j],
                       batch.motif_score_matrix(j, L))

    for i in my_ids:
        if errors[i] is not None:
            submit(i, error=errors[i])
    try:
        if good:
            run(good)
    except Exception:  # noqa: BLE001
        # second attempt, motif by motif (orientation pairs together): what still fails is recorded for that motif only
        first = traceback.format_exc()
        done = set(submitted)
        units, prev = [], None
        for i in good:
            if i in done:
                continue
            key = tups[i][0]
            if key == prev:
                units[-1].append(i)
            else:
                units.append([i])
            prev = key
        for unit in units:
            for i in unit:
                attempts[i] = 2
            if stop_max_attempt_number is not None and stop_max_attempt_number < 2:
                for i in unit:
                    attempts[i] = 1
                    submit(i, error=first)
                continue
            try:
                run(unit)
            except Exception:  # noqa: BLE001
                err = traceback.format_exc()
                for i in unit:
                    if i not in set(submitted):
                        submit(i, error=err)
    results = tw.results()
    if own_writer:
        tw.close()
    rows = {}
    for i, (retval, yaml_fp, log_fp) in zip(submitted, results):
        motif_id, orient, outdir = tups[i]
        rows[i] = (motif_id, orient, outdir, yaml_fp or normpath(f'{outdir}/wrapped_params.yaml'),
                   log_fp or normpath(f'{outdir}/wrapped_params.log'), retval, attempts[i])

    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, rows)
        rows = {}
        for g in gathered:
            rows.update(g)
    filepaths_df = pd.DataFrame([rows[i] for i in sorted(rows)],
                                columns=['motif_id', 'orientation', 'outdir', 'params', 'log', 'retval', 'attempts'])
    if world > 1:
        dist.barrier()
    if keep_dataset is False and rank == 0:
        shutil.rmtree(normpath(dataset_filepath))            # batch.py:187-188 (the .element_spec.p sidecar stays)
    if free_dataset:
        dataset.release()
    return filepaths_df

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
from .single import _task_filepaths, write_task_outputs

orientation_to_filepath = {'+': 'fwd', '-': 'rev', '+/-': 'fwd-rev'}


def slugify(text):
    """python-slugify's default behaviour for the ASCII motif ids MEPP uses (the package is not a
    dependency here): NFKD fold to ASCII, lower-case, every run of non-alphanumerics -> '-'."""
    text = unicodedata.normalize('NFKD', str(text)).encode('ascii', 'ignore').decode('ascii')
    text = re.sub(r"[']+", '', text.lower())
    return re.sub(r'[^a-z0-9]+', '-', text).strip('-')


motif_id_and_orientation_to_filepath = lambda motif_id, orientation: (
    normpath(f'{slugify(motif_id)}/orientation_{orientation_to_filepath[orientation]}'))


def run_batch(
    dataset, motif_matrix_dict, out_filepath, center=None, motif_orientations=['+', '-', '+/-'], motif_margin=1,
    motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None, confidence_interval_pct=95, motif_score_sigma=1.0,
    motif_score_cmap='gray', rank_smoothing_factor=5, figure_width=10, figure_height=10, save_datasets=False,
    save_profile_data=True, save_motif_score_matrix=False, save_figures=True, figure_formats=['png', 'svg'],
    figure_dpi=300, n_jobs=1, free_dataset=True, keep_dataset=False, progress_wrapper=None, no_gpu=False,
    stop_max_attempt_number=3, wait_random_min=1.0, wait_random_max=2.0
):
    if no_gpu:
        from .utils import force_cpu_only
        force_cpu_only()
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    dataset_filepath = normpath(f'{out_filepath}/dataset')
    if rank == 0:
        os.makedirs(dataset_filepath, exist_ok=True)
        save_dataset(dataset, dataset_filepath, delete_existing=True)      # batch.py:60-70 (kept for --keepdata)

    tups = [(motif_id, orient,
             normpath(f'{out_filepath}/' + motif_id_and_orientation_to_filepath(motif_id, orient)))
            for motif_id in motif_matrix_dict for orient in motif_orientations]      # batch.py:73-90 order
    tasks = [(motif_matrix_dict[mid], o) for mid, o, _ in tups]
    my_ids = parallel.shard_task_units(tasks, world)[rank]      # orientation pairs of a motif stay on one rank

    rows = {}
    batch = None
    err = None
    try:
        batch = profile_tasks(dataset, [tasks[i] for i in my_ids], motif_margin=motif_margin,
                              confidence_interval_pct=confidence_interval_pct, center=center,
                              return_hits=save_motif_score_matrix)
    except Exception:  # noqa: BLE001 -- recorded per task below, like a failed child process
        err = traceback.format_exc()
    for j, i in enumerate(my_ids):
        motif_id, orient, outdir = tups[i]
        os.makedirs(outdir, exist_ok=True)
        motif_matrix_filepath = normpath(f'{outdir}/motif_matrix.npy')
        np.save(motif_matrix_filepath, motif_matrix_dict[motif_id])
        params = dict(
            dataset_filepath=dataset_filepath, motif_matrix_filepath=motif_matrix_filepath, motif_id=motif_id,
            output_filepath=outdir, center=center, motif_orientation=orient, motif_margin=motif_margin,
            motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=bg,
            confidence_interval_pct=confidence_interval_pct, motif_score_sigma=motif_score_sigma,
            motif_score_cmap=motif_score_cmap, rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width,
            figure_height=figure_height, save_datasets=save_datasets, save_profile_data=save_profile_data,
            save_motif_score_matrix=save_motif_score_matrix, save_figures=True, figure_formats=list(figure_formats),
            figure_dpi=figure_dpi, no_gpu=no_gpu)
        yaml_fp, log_fp = normpath(f'{outdir}/wrapped_params.yaml'), normpath(f'{outdir}/wrapped_params.log')
        import yaml
        with open(yaml_fp, 'w', encoding='utf8') as f:
            yaml.dump(params, f, default_flow_style=False, allow_unicode=True)
        retval = 0
        with open(log_fp, 'w') as lf:
            if batch is None:
                lf.write(err or 'task failed')
                retval = 1
            else:
                try:
                    fps = _task_filepaths(outdir, dataset_filepath, motif_matrix_filepath, figure_formats)
                    fps.pop('mepp_plot', None)                 # figures: plotting layer, out of scope
                    msm = batch.motif_score_matrix(j, dataset.sequence_length) if save_motif_score_matrix else None
                    run_params = {k: v for k, v in params.items() if k not in ('center', 'no_gpu')}
                    write_task_outputs(run_params, fps, msm, batch.positions_df(j), batch.ranks_df(j),
                                       save_profile_data, save_motif_score_matrix)
                except Exception:  # noqa: BLE001
                    traceback.print_exc(file=lf)
                    retval = 1
        rows[i] = (motif_id, orient, outdir, yaml_fp, log_fp, retval, 1)

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

"""Per-task driver with the reference's names and file layout (reference: mepp/single.py).

run_single keeps mepp/single.py:57-302's keyword arguments and writes the same per-task files
(positions_df.{tsv,pkl}, ranks_df.{tsv,pkl}, params.yaml, filepaths.yaml, optional
motif_score_matrix.npz); wrap_single keeps mepp/single.py:308-408's signature and return value
(yaml_filepath, log_filepath, retval, attempts) but runs the task IN PROCESS instead of spawning
`python -m mepp.single` (the subprocess existed to work around TensorFlow + multiprocessing,
single.py:48-49).  Figures (mepp_plot.*) belong to the reference's plotting layer, which is out of
this engine's scope; they are produced only if a `plot` module with plot_profile_data is importable.
"""
import io
import os
import sys
import traceback
from os.path import normpath

import numpy as np
import yaml

from .io import load_dataset
from .utils import force_cpu_only, manage_gpu_memory
from .core import dataset_and_motif_matrix_to_profile_data


def _task_filepaths(output_filepath, dataset_filepath, motif_matrix_filepath, figure_formats):
    """single.py:111-170."""
    fp = dict(
        filepaths=f'{output_filepath}/filepaths.yaml', params=f'{output_filepath}/params.yaml',
        output=output_filepath, dataset=dataset_filepath, motif_matrix=motif_matrix_filepath,
        processed_dataset=f'{output_filepath}/processed_dataset',
        smoothed_processed_dataset=f'{output_filepath}/smoothed_processed_dataset',
        motif_score_matrix=f'{output_filepath}/motif_score_matrix.npz')
    fp = {k: normpath(v) if v is not None else None for k, v in fp.items()}
    fp['positions_df'] = dict(tsv=normpath(f'{output_filepath}/positions_df.tsv'),
                              pickle=normpath(f'{output_filepath}/positions_df.pkl'))
    fp['ranks_df'] = dict(tsv=normpath(f'{output_filepath}/ranks_df.tsv'),
                          pickle=normpath(f'{output_filepath}/ranks_df.pkl'))
    fp['mepp_plot'] = {fmt: normpath(f'{output_filepath}/mepp_plot.{fmt}') for fmt in figure_formats}
    return fp


def write_task_outputs(params, filepaths, motif_score_matrix, positions_df, ranks_df, save_profile_data=True,
                       save_motif_score_matrix=False):
    """The file-level ABI of one task directory (single.py:198-302)."""
    os.makedirs(filepaths['output'], exist_ok=True)
    filepaths = dict(filepaths)
    filepaths.pop('processed_dataset', None)          # save_datasets (TF snapshots of lazy datasets) is not supported
    filepaths.pop('smoothed_processed_dataset', None)
    if save_profile_data:
        if save_motif_score_matrix:
            np.savez_compressed(filepaths['motif_score_matrix'], motif_score_matrix=np.asarray(motif_score_matrix))
        else:
            filepaths.pop('motif_score_matrix', None)
        positions_df.to_pickle(filepaths['positions_df']['pickle'])
        positions_df.to_csv(filepaths['positions_df']['tsv'], sep='\t', index=False)
        ranks_df.to_pickle(filepaths['ranks_df']['pickle'])
        ranks_df.to_csv(filepaths['ranks_df']['tsv'], sep='\t', index=False)
    else:
        for k in ('motif_score_matrix', 'positions_df', 'ranks_df'):
            filepaths.pop(k, None)
    with io.open(filepaths['params'], 'w', encoding='utf8') as f:
        yaml.dump(params, f, default_flow_style=False, allow_unicode=True)
    with io.open(filepaths['filepaths'], 'w', encoding='utf8') as f:
        yaml.dump(filepaths, f, default_flow_style=False, allow_unicode=True)
    return filepaths


def _maybe_plot(filepaths, motif_score_matrix, positions_df, ranks_df, motif_matrix, params):
    try:
        from .plot import plot_profile_data  # optional, not part of the engine
    except ImportError:
        filepaths.pop('mepp_plot', None)
        return
    import matplotlib.pyplot as plt
    stride = -1 if params['motif_orientation'] == '-' else 1
    fig, _ = plot_profile_data(motif_score_matrix, positions_df, ranks_df, motif_matrix=motif_matrix[::stride, ::stride],
                               title=(f"Motif enrichment positional profile for \n {params['motif_id']} \n "
                                      f"Orientation: {params['motif_orientation']}"),
                               figsize=(params['figure_width'], params['figure_height']),
                               rank_smoothing_factor=params['rank_smoothing_factor'],
                               motif_score_sigma=params['motif_score_sigma'], motif_score_cmap=params['motif_score_cmap'])
    for fmt in params['figure_formats']:
        fig.savefig(filepaths['mepp_plot'][fmt], dpi=params['figure_dpi'])
    plt.close(fig)


def run_single(
    dataset_filepath=None, motif_matrix_filepath=None, motif_id=None, output_filepath=None, center=None,
    motif_orientation='+', motif_margin=0, motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None,
    confidence_interval_pct=95, motif_score_sigma=1.0, motif_score_cmap='gray', rank_smoothing_factor=5,
    figure_width=10, figure_height=10, save_datasets=False, save_profile_data=True, save_motif_score_matrix=False,
    save_figures=True, figure_formats=['png', 'svg'], figure_dpi=300, no_gpu=False, dataset=None, dp_precision=None
):
    """single.py:57-302.  dp_precision (extension): grid of the MOODS threshold DP, recorded in params.yaml.  `dataset` (an already loaded ScoredDataset) is an extension that lets the batch
    driver skip the per-task load_dataset of single.py:177."""
    if no_gpu:
        force_cpu_only()
    else:
        manage_gpu_memory()
    params = dict(
        dataset_filepath=dataset_filepath, motif_matrix_filepath=motif_matrix_filepath, motif_id=motif_id,
        output_filepath=output_filepath, motif_orientation=motif_orientation, motif_margin=motif_margin,
        motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=bg,
        confidence_interval_pct=confidence_interval_pct, motif_score_sigma=motif_score_sigma,
        motif_score_cmap=motif_score_cmap, rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width,
        figure_height=figure_height, save_datasets=save_datasets, save_profile_data=save_profile_data,
        save_motif_score_matrix=save_motif_score_matrix, save_figures=save_figures, figure_formats=figure_formats,
        figure_dpi=figure_dpi)
    from .motif_layers import DP_PRECISION
    params['dp_precision'] = float(DP_PRECISION if dp_precision is None else dp_precision)
    filepaths = _task_filepaths(output_filepath, dataset_filepath, motif_matrix_filepath, figure_formats)
    os.makedirs(filepaths['output'], exist_ok=True)
    if dataset is None:
        dataset = load_dataset(filepaths['dataset'])
    motif_matrix = np.load(filepaths['motif_matrix'])
    (motif_score_matrix, positions_df, ranks_df, _processed, _smoothed) = dataset_and_motif_matrix_to_profile_data(
        dataset, motif_matrix, orientation=motif_orientation, motif_margin=motif_margin,
        motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=bg, center=center,
        confidence_interval_pct=confidence_interval_pct, dp_precision=dp_precision)
    if save_figures:
        _maybe_plot(filepaths, motif_score_matrix, positions_df, ranks_df, motif_matrix, params)
    else:
        filepaths.pop('mepp_plot', None)
    write_task_outputs(params, filepaths, motif_score_matrix, positions_df, ranks_df, save_profile_data,
                       save_motif_score_matrix)


def run_single_from_yaml(yaml_filepath):
    """single.py:37-55 (`python -m mepp_b200.single -y params.yaml`)."""
    with open(yaml_filepath, 'r') as yaml_file:
        params = yaml.safe_load(yaml_file)
    params.pop('yaml_filepath', None)
    return run_single(**params)


def generate_subprocess_args(yaml_filepath):
    """single.py:304-306."""
    return ['python', '-m', __name__, '-y', yaml_filepath]


def wrap_single(
    dataset_filepath, motif_matrix, motif_id, output_filepath, center=None, motif_orientation='+', motif_margin=0,
    motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None, confidence_interval_pct=95, motif_score_sigma=1.0,
    motif_score_cmap='gray', rank_smoothing_factor=5, figure_width=10, figure_height=10, save_datasets=True,
    save_profile_data=True, save_motif_score_matrix=False, save_figures=True, figure_formats=['png', 'svg'],
    figure_dpi=300, no_gpu=False, subprocess_args_func=generate_subprocess_args, stop_max_attempt_number=3,
    wait_random_min=1.0, wait_random_max=2.0, dataset=None
):
    """single.py:308-408: writes motif_matrix.npy + wrapped_params.yaml, runs the task, returns
    (yaml_filepath, log_filepath, retval, attempts).  The task runs in process; an exception becomes
    retval 1 with the traceback in wrapped_params.log (the reference's child stderr), never raised."""
    motif_matrix_filepath = normpath(f'{output_filepath}/motif_matrix.npy')
    params = dict(
        dataset_filepath=dataset_filepath, motif_matrix_filepath=motif_matrix_filepath, motif_id=motif_id,
        output_filepath=output_filepath, center=center, motif_orientation=motif_orientation, motif_margin=motif_margin,
        motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=bg,
        confidence_interval_pct=confidence_interval_pct, motif_score_sigma=motif_score_sigma,
        motif_score_cmap=motif_score_cmap, rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width,
        figure_height=figure_height, save_datasets=save_datasets, save_profile_data=save_profile_data,
        save_motif_score_matrix=save_motif_score_matrix, save_figures=save_figures, figure_formats=figure_formats,
        figure_dpi=figure_dpi, no_gpu=no_gpu)
    os.makedirs(normpath(output_filepath), exist_ok=True)
    np.save(motif_matrix_filepath, motif_matrix)
    yaml_filepath = normpath(f'{output_filepath}/wrapped_params.yaml')
    log_filepath = normpath(f'{output_filepath}/wrapped_params.log')
    with io.open(yaml_filepath, 'w', encoding='utf8') as yaml_file:
        yaml.dump(params, yaml_file, default_flow_style=False, allow_unicode=True)
    retval, attempts = 0, 1
    with open(log_filepath, 'w') as lf:
        try:
            run_single(dataset=dataset, **params)
        except Exception:  # noqa: BLE001 -- the reference records failures, it never raises (single.py:392-408)
            traceback.print_exc(file=lf)
            retval = 1
    return yaml_filepath, log_filepath, retval, attempts


def _main(argv=None):
    import click

    @click.command()
    @click.option('-y', '--yaml', 'yaml_filepath', required=True, type=str,
                  help='Filepath to a YAML file describing the parameters for this MEPP.')
    def cmd(yaml_filepath):
        run_single_from_yaml(yaml_filepath)
    return cmd(args=argv, standalone_mode=False)


if __name__ == '__main__':
    sys.exit(_main())

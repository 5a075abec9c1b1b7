"""End-to-end orchestration (reference: mepp/mepp.py:47-299): scored FASTA + motif file -> staged
dataset -> run_batch -> filepaths.tsv.  Same signature and return triple as the reference's run_mepp.
The reference's clustering / HTML reports (mepp.py:269-297) consume the per-task files written here
and are outside this engine's scope: the two returned dicts are empty unless an optional
`clustering` module is importable."""
import multiprocessing
import os
from os.path import normpath

from .batch import run_batch, orientation_to_filepath
from .io import motif_matrix_filepath_to_dicts, scored_fasta_filepath_to_dicts
from .permutations import add_permuted_scores_to_dataset
from .utils import (append_gc_ratios_to_dataset, filter_scored_fasta_df, force_cpu_only, manage_gpu_memory,
                    order_scored_fasta_df, scored_fasta_df_to_dataset, scored_fasta_dicts_to_df)


LAST_RUN_TIMINGS = {}      # wall-clock phases of the last run_mepp call of this process (bench.py --mode fullrun)


def run_mepp(
    scored_fasta_filepath, motifs_filepath, out_filepath,
    center=None, degenerate_pct_thresh=100, num_permutations=1000, batch_size=1000,
    n_cpu_jobs=multiprocessing.cpu_count(),
    motif_orientations=['+', '+/-'], motif_margin=2, motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None,
    confidence_interval_pct=95, motif_score_sigma=0.5, motif_score_cmap='gray', rank_smoothing_factor=5,
    figure_width=10, figure_height=10, figure_formats=['png', 'svg'], figure_dpi=300, n_gpu_jobs=3,
    save_datasets=False, save_profile_data=True, save_motif_score_matrix=False, keep_dataset=False,
    stop_max_attempt_number=3, wait_random_min=1.0, wait_random_max=2.0,
    cluster_method='average', cluster_metric='correlation', mepp_plot_format='png', table_image_dpi=100,
    table_image_format='png', mt_method='fdr_tsbky', mt_alpha=0.01, thorough_mt=True, no_gpu=False,
    permutation_seed=None, shuffle_seed=None, dp_precision=None
):
    gc_ratio_type = 'global'                                               # mepp.py:96
    import time
    import numpy as np
    import torch.distributed as dist
    timings = LAST_RUN_TIMINGS
    timings.clear()
    t_run = time.perf_counter()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    os.makedirs(normpath(out_filepath), exist_ok=True)
    if no_gpu:
        force_cpu_only()
    else:
        manage_gpu_memory()
    # the writer processes of the task directories start now, so that their interpreter start-up (numpy, pandas)
    # is hidden behind the parsing of the inputs
    from .writer import TaskWriter
    try:
        big = os.path.getsize(scored_fasta_filepath) > (16 << 20)
    except (OSError, TypeError):
        big = False
    task_writer = TaskWriter(None, prestart=big)
    try:
        if multi:
            # every rank parses the same files; the shuffle and permutation seeds (unseeded in the reference:
            # utils.py:117, permutations.py:33-37) are drawn ONCE, on rank 0, so that all ranks order the rows and
            # permute the scores identically; only rank 0 writes the run-level files
            seeds = [shuffle_seed if shuffle_seed is not None else int(np.random.SeedSequence().entropy % (1 << 32)),
                     permutation_seed if permutation_seed is not None else
                     int(np.random.SeedSequence().entropy % (1 << 63))]
            dist.broadcast_object_list(seeds, src=0)
            shuffle_seed, permutation_seed = seeds
        sequence_dict, score_dict, description_dict = scored_fasta_filepath_to_dicts(scored_fasta_filepath)
        scored_fasta_df = order_scored_fasta_df(
            filter_scored_fasta_df(scored_fasta_dicts_to_df(sequence_dict, score_dict, description_dict),
                                   degenerate_pct_thresh=degenerate_pct_thresh),
            random_state=shuffle_seed)
        timings['parse_filter_order_s'] = time.perf_counter() - t_run
        # filtered_data_df.{tsv,pkl} (mepp.py:99-104; 100 MB of text at cfg 3) are written by a thread beside the
        # staging of the dataset and the tasks
        from concurrent.futures import ThreadPoolExecutor
        side = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mepp-filtered-df")

        def write_filtered():
            scored_fasta_df.to_csv(normpath(f'{out_filepath}/filtered_data_df.tsv'), sep='\t', index=False)
            scored_fasta_df.to_pickle(normpath(f'{out_filepath}/filtered_data_df.pkl'))

        filtered_written = side.submit(write_filtered) if rank == 0 else None
        t0 = time.perf_counter()
        dataset = add_permuted_scores_to_dataset(
            append_gc_ratios_to_dataset(scored_fasta_df_to_dataset(scored_fasta_df, batch_size=batch_size,
                                                                   n_jobs=n_cpu_jobs), gc_ratio_type=gc_ratio_type),
            batch_size=batch_size, num_permutations=num_permutations, seed=permutation_seed)
        task_writer.set_scores(dataset.scores)
        motif_matrix_dict, _motif_consensus_dict = motif_matrix_filepath_to_dicts(motifs_filepath)
        timings['dataset_permutations_motifs_s'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        filepaths_df = run_batch(
            dataset=dataset, motif_matrix_dict=motif_matrix_dict, out_filepath=out_filepath, center=center,
            motif_orientations=motif_orientations, motif_margin=motif_margin, motif_pseudocount=motif_pseudocount,
            motif_pvalue=motif_pvalue, bg=bg, confidence_interval_pct=confidence_interval_pct,
            motif_score_sigma=motif_score_sigma, motif_score_cmap=motif_score_cmap,
            rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width, figure_height=figure_height,
            save_datasets=save_datasets, keep_dataset=keep_dataset, save_profile_data=save_profile_data,
            save_motif_score_matrix=save_motif_score_matrix, save_figures=True, figure_formats=figure_formats,
            figure_dpi=figure_dpi, n_jobs=n_gpu_jobs, no_gpu=no_gpu, stop_max_attempt_number=stop_max_attempt_number,
            wait_random_min=wait_random_min, wait_random_max=wait_random_max, task_writer=task_writer,
            dp_precision=dp_precision)
        timings['run_batch_s'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if filtered_written is not None:
            filtered_written.result()
        side.shutdown(wait=True)
        timings['wait_filtered_df_s'] = time.perf_counter() - t0
    finally:
        task_writer.close()
    if rank == 0:
        filepaths_df.to_csv(normpath(f'{out_filepath}/filepaths.tsv'), sep='\t', index=False)     # mepp.py:197-199
    timings['total_s'] = time.perf_counter() - t_run
    results_html_filepaths, clustermap_html_filepaths = {}, {}
    try:
        from .clustering import filepaths_df_to_clustering_results  # optional reporting layer (out of scope)
    except ImportError:
        # tabular part of the results tables (no images / HTML): mepp.py:285-297 file names
        from .summary import results_table
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:
            for motif_orientation in motif_orientations:
                o = orientation_to_filepath[motif_orientation]
                table = results_table(filepaths_df, motif_orientation, mt_method=mt_method, mt_alpha=mt_alpha,
                                      thorough_mt=thorough_mt)
                table.to_csv(normpath(f'{out_filepath}/results_table_orientation_{o}.tsv'), sep='\t', index=False)
                table.to_pickle(normpath(f'{out_filepath}/results_table_orientation_{o}.pkl'))
        return filepaths_df, results_html_filepaths, clustermap_html_filepaths
    for motif_orientation in motif_orientations:
        o = orientation_to_filepath[motif_orientation]
        res = filepaths_df_to_clustering_results(
            out_filepath, center=center, motif_orientation=motif_orientation, cluster_method=cluster_method,
            cluster_metric=cluster_metric, mepp_plot_format=mepp_plot_format, mt_method=mt_method, mt_alpha=mt_alpha,
            thorough_mt=thorough_mt, table_image_dpi=table_image_dpi, table_image_format=table_image_format,
            n_jobs=n_cpu_jobs)
        res['results_df'].to_csv(normpath(f'{out_filepath}/results_table_orientation_{o}.tsv'), sep='\t', index=False)
        res['results_df'].to_pickle(normpath(f'{out_filepath}/results_table_orientation_{o}.pkl'))
        for kind, key, store in (('results_table', 'results_html', results_html_filepaths),
                                 ('clustermap', 'clustermap_html', clustermap_html_filepaths)):
            path = normpath(f'{out_filepath}/{kind}_orientation_{o}.html')
            with open(path, 'w') as f:
                f.write(res[key])
            store[motif_orientation] = path
    return filepaths_df, results_html_filepaths, clustermap_html_filepaths

"""`mepp` command line (reference: mepp/cli.py:11-578): the same flags, forwarded to run_mepp.

Flags that configure the reference's plotting / clustering / HTML layers are accepted for drop-in
compatibility and passed through; --nogpu is an error (this engine has no CPU fallback); --gjobs /
--attempts / --minwait / --maxwait no longer matter (tasks run in process, no retries needed).
"""
import multiprocessing
import sys

import click

from .mepp import run_mepp

_CMETHODS = ['average', 'single', 'complete', 'centroid', 'median', 'ward', 'weighted']
_CMETRICS = ['correlation', 'euclidean', 'minkowski', 'cityblock', 'seuclidean', 'sqeuclidean', 'cosine',
             'jensenshannon', 'chebyshev', 'canberra', 'braycurtis', 'mahalanobis']
_MTMETHODS = ['fdr_tsbky', 'bonferroni', 'sidak', 'holm-sidak', 'holm', 'simes-hochberg', 'hommel', 'fdr_bh',
              'fdr_by', 'fdr_tsbh']


@click.command()
@click.option('--fa', 'scored_fasta_filepath', type=str, required=True,
              help='Path to a scored fasta file (">id score" headers), "-" for stdin.')
@click.option('--motifs', 'motifs_filepath', type=str, required=True, help='Path to a JASPAR-format motif file.')
@click.option('--out', 'out_filepath', type=str, required=True, help='Output directory.')
@click.option('--center', 'center', type=int, default=None, help='0-based center position. Default: L // 2.')
@click.option('--dgt', 'degenerate_pct_thresh', type=float, default=100.0,
              help='Drop sequences with more than this percentage of degenerate (non upper-case ACGT) bases.')
@click.option('--perms', 'num_permutations', type=int, default=1000, help='Number of score permutations.')
@click.option('--batch', 'batch_size', type=int, default=1000, help='Kept for compatibility (no effect).')
@click.option('--jobs', 'n_cpu_jobs', type=int, default=multiprocessing.cpu_count(), help='Host threads.')
@click.option('--keepdata', 'keep_dataset', is_flag=True, help='Keep the staged dataset directory.')
@click.option('--orientations', 'motif_orientations', type=str, default='+,+/-',
              help='Comma-separated motif orientations out of +,-,+/-.')
@click.option('--margin', 'motif_margin', type=int, default=2, help='Blur motif matches by +/- this many bp.')
@click.option('--pcount', 'motif_pseudocount', type=float, default=0.0001, help='Motif pseudocount (see README).')
@click.option('--pval', 'motif_pvalue', type=float, default=0.0001, help='Motif match p-value (see README).')
@click.option('--bg', 'bg', nargs=4, type=float, default=(0.25, 0.25, 0.25, 0.25), help='Background A C G T.')
@click.option('--ci', 'confidence_interval_pct', type=float, default=95.0, help='Confidence interval percent.')
@click.option('--sigma', 'motif_score_sigma', type=float, default=0.5)
@click.option('--cmap', 'motif_score_cmap', type=str, default='gray_r')
@click.option('--smoothing', 'rank_smoothing_factor', type=int, default=5)
@click.option('--width', 'figure_width', type=int, default=10)
@click.option('--height', 'figure_height', type=int, default=10)
@click.option('--formats', 'figure_formats', type=str, default='png,svg')
@click.option('--dpi', 'figure_dpi', type=int, default=300)
@click.option('--gjobs', 'n_gpu_jobs', type=int, default=1, help='Kept for compatibility (no effect).')
@click.option('--nogpu', 'no_gpu', is_flag=True, help='Not supported: this engine has no CPU fallback.')
@click.option('--attempts', 'stop_max_attempt_number', type=int, default=10)
@click.option('--minwait', 'wait_random_min', type=float, default=1.0)
@click.option('--maxwait', 'wait_random_max', type=float, default=1.0)
@click.option('--cmethod', 'cluster_method', type=click.Choice(_CMETHODS, case_sensitive=False), default='average')
@click.option('--cmetric', 'cluster_metric', type=click.Choice(_CMETRICS, case_sensitive=False), default='correlation')
@click.option('--tdpi', 'table_image_dpi', type=int, default=100)
@click.option('--tformat', 'table_image_format', type=click.Choice(['png', 'svg'], case_sensitive=False), default='png')
@click.option('--mtmethod', 'mt_method', type=click.Choice(_MTMETHODS, case_sensitive=False), default='fdr_by')
@click.option('--mtalpha', 'mt_alpha', type=float, default=0.01)
@click.option('--thoroughmt', 'thorough_mt', flag_value=True, default=True)
@click.option('--non-thoroughmt', 'thorough_mt', flag_value=False, default=True)
def main(scored_fasta_filepath, motifs_filepath, out_filepath, center=None, degenerate_pct_thresh=100.0,
         num_permutations=1000, batch_size=1000, n_cpu_jobs=multiprocessing.cpu_count(), keep_dataset=False,
         motif_orientations='+,+/-', motif_margin=2, motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None,
         confidence_interval_pct=95.0, motif_score_sigma=0.5, motif_score_cmap='gray', rank_smoothing_factor=5,
         figure_width=10, figure_height=10, figure_formats='png,svg', figure_dpi=300, n_gpu_jobs=3, no_gpu=False,
         stop_max_attempt_number=10, wait_random_min=1.0, wait_random_max=2.0, cluster_method='average',
         cluster_metric='correlation', table_image_dpi=100, table_image_format='png', mt_method='fdr_tsbky',
         mt_alpha=0.01, thorough_mt=True):
    """Profile positional enrichment of motifs in a list of scored sequences (MEPP), on a B200."""
    if no_gpu:
        raise click.UsageError('--nogpu is not supported: mepp_b200 has no CPU fallback')
    orientations = [x.strip() for x in motif_orientations.split(',')]           # cli.py:503-507
    for o in orientations:
        if o not in ('+', '-', '+/-'):
            raise click.UsageError(f'unknown orientation {o!r}')
    formats = [x.strip().lower() for x in figure_formats.split(',')]
    filepaths_df, results_html, clustermap_html = run_mepp(
        scored_fasta_filepath, motifs_filepath, out_filepath, center=center,
        degenerate_pct_thresh=degenerate_pct_thresh, num_permutations=num_permutations, batch_size=batch_size,
        n_cpu_jobs=n_cpu_jobs, keep_dataset=keep_dataset, motif_orientations=orientations, motif_margin=motif_margin,
        motif_pseudocount=motif_pseudocount, motif_pvalue=motif_pvalue, bg=list(bg) if bg is not None else None,
        confidence_interval_pct=confidence_interval_pct, motif_score_sigma=motif_score_sigma,
        motif_score_cmap=motif_score_cmap, rank_smoothing_factor=rank_smoothing_factor, figure_width=figure_width,
        figure_height=figure_height, figure_formats=formats, figure_dpi=figure_dpi, n_gpu_jobs=n_gpu_jobs,
        no_gpu=no_gpu, stop_max_attempt_number=stop_max_attempt_number, wait_random_min=wait_random_min,
        wait_random_max=wait_random_max, cluster_method=cluster_method.lower(), cluster_metric=cluster_metric.lower(),
        mepp_plot_format=formats[0], table_image_dpi=table_image_dpi, table_image_format=table_image_format.lower(),
        mt_method=mt_method.lower(), mt_alpha=mt_alpha, thorough_mt=thorough_mt)
    click.echo(f'{len(filepaths_df)} tasks written under {out_filepath} '
               f'({int((filepaths_df["retval"] != 0).sum())} failed)')
    for k, v in results_html.items():
        click.echo(f'HTML result table for orientation {k}: {v}')
    for k, v in clustermap_html.items():
        click.echo(f'HTML clustermap for orientation {k}: {v}')
    return 0


if __name__ == "__main__":
    sys.exit(main())  # pragma: no cover

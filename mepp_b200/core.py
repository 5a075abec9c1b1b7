"""The profiling hot path behind the reference's entry point (reference: mepp/core.py:490-632).

dataset_and_motif_matrix_to_profile_data keeps the reference's signature and 5-tuple return;
the work happens in Engine (csrc/*.cu through the C ABI).  profile_tasks is the batched form
the in-process task scheduler (batch.py) uses: many (motif, orientation) tasks per call.
"""
import numpy as np

from . import stats_host
from .utils import get_sequence_length_from_dataset

# re-exported host helpers with the reference's names (core.py:380-413, :438-454)
POSITIONS_COLUMNS = stats_host.POSITIONS_COLUMNS


class ProcessedDataset:
    """Stands in for the lazily mapped tf.data datasets the reference returns as the 4th / 5th
    tuple element (core.py:540-552, :626-632): materialises the (N, L) score matrix on demand."""

    def __init__(self, motif_score_matrix, scores, margin=0):
        self._x0 = motif_score_matrix
        self.scores = scores
        self.margin = margin

    def motif_scores(self):
        X0 = np.asarray(self._x0)
        mg = self.margin
        if mg <= 0:
            return X0
        # core.py:59-145 blur on the host, only when someone asks for the smoothed dataset
        N, L = X0.shape
        k = 1 + 2 * mg
        out = np.zeros_like(X0)
        if L >= k:
            acc = X0[:, 0:L - k + 1].copy()
            for d in range(1, k):
                acc = acc + X0[:, d:L - k + 1 + d]
            out[:, mg:L - mg] = acc / np.float32(k)
        return out


def profile_tasks(dataset, tasks, motif_margin=0, confidence_interval_pct=95, center=None, return_hits=False,
                  return_r=False, honour_flags=False, motif_pseudocount=0.0001, motif_pvalue=0.0001, bg=None,
                  on_group=None, dp_precision=None):
    """Batched hot path: tasks = [(motif_matrix (4, m), orientation), ...] -> engine.ProfileBatch.
    on_group(t0, t1, positions, density) is called when the results of tasks t0 .. t1-1 are on the host."""
    eng = dataset.engine()
    return eng.profile(tasks, margin=motif_margin, confidence_interval_pct=confidence_interval_pct, center=center,
                       return_hits=return_hits, return_r=return_r, honour_flags=honour_flags,
                       pseudocount=motif_pseudocount, pvalue=motif_pvalue, bg=bg, on_group=on_group,
                       dp_precision=dp_precision)


def dataset_and_motif_matrix_to_profile_data(
    dataset,
    motif_matrix,
    orientation='+',
    motif_margin=0,
    motif_pseudocount=0.0001,
    motif_pvalue=0.0001,
    bg=None,
    confidence_interval_pct=95,
    center=None,
    dp_precision=None
):
    """Same signature and return as core.py:490-632 (dp_precision: extension, grid of the MOODS threshold DP):
    (motif_score_matrix, positions_df, ranks_df, processed_dataset, smoothed_processed_dataset).

    As in the reference, motif_pseudocount / motif_pvalue / bg are accepted but do not reach the
    threshold computation (core.py:513-534 never forwards them); orientation '+' / '-' select a
    single strand, anything else scans both strands (core.py:511-534).
    """
    sequence_length = get_sequence_length_from_dataset(dataset)
    batch = profile_tasks(dataset, [(motif_matrix, orientation)], motif_margin=motif_margin,
                          confidence_interval_pct=confidence_interval_pct, center=center, return_hits=True,
                          dp_precision=dp_precision)
    motif_score_matrix = batch.motif_score_matrix(0, sequence_length)
    positions_df = batch.positions_df(0)
    ranks_df = batch.ranks_df(0)
    processed = ProcessedDataset(motif_score_matrix, batch.scores, 0)
    smoothed = ProcessedDataset(motif_score_matrix, batch.scores, motif_margin) if motif_margin > 0 else processed
    return motif_score_matrix, positions_df, ranks_df, processed, smoothed

"""Permuted-score columns (reference: mepp/permutations.py:5-58).

The reference builds an (N, 1+P) float32 matrix whose column 0 is the true score and whose
other columns are tf.random.shuffle(scores) -- unseeded (permutations.py:33-37).  Here the
permutations are an explicit (P, N) int32 index matrix generated on the host (optionally
seeded) so that the CPU oracle and the GPU path see identical permutations; the (N, 1+P)
matrix itself is only ever materialised on the GPU, as the tiled fp16 GEMM operand.
"""
import numpy as np


def make_permutation_indices(num_sequences, num_permutations, seed=None):
    """(P, N) int32; row p is a uniform random permutation (numpy Generator PCG64(seed))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((num_permutations, num_sequences), dtype=np.int32)
    base = np.arange(num_sequences, dtype=np.int32)
    for p in range(num_permutations):
        out[p] = rng.permutation(base)
    return out


def add_permuted_scores_to_dataset(dataset, batch_size=None, num_permutations=1000, perm_indices=None, seed=None):
    """Same call shape as the reference (permutations.py:5-9); P <= 0 returns the dataset unchanged (:10-11)."""
    if perm_indices is None and num_permutations <= 0:
        return dataset
    if perm_indices is None:
        perm_indices = make_permutation_indices(dataset.num_sequences, num_permutations, seed)
    perm_indices = np.asarray(perm_indices, dtype=np.int32)
    if perm_indices.ndim != 2 or perm_indices.shape[1] != dataset.num_sequences:
        raise ValueError("perm_indices must be (P, N)")
    kw = dict(perm_idx=perm_indices)
    if batch_size is not None:
        kw['batch_size'] = batch_size
    return dataset.replace(**kw)

"""Permuted-score columns (reference: mepp/permutations.py:5-58).

The reference builds an (N, 1+P) float32 matrix whose column 0 is the true score and whose
other columns are tf.random.shuffle(scores) -- unseeded (permutations.py:33-37).  Here the
permutations are an explicit (P, N) int32 index matrix generated on the host (optionally
seeded) so that the CPU oracle and the GPU path see identical permutations; the (N, 1+P)
matrix itself is only ever materialised on the GPU, as the tiled fp16 GEMM operand.
"""
import numpy as np


PERM_CHUNK = 64     # rows per independent random stream: the result does not depend on the number of threads


def make_permutation_indices(num_sequences, num_permutations, seed=None, threads=None, out=None):
    """(P, N) int32; row p is a uniform random permutation.  Rows are drawn in chunks of PERM_CHUNK, chunk c from
    numpy Generator(PCG64(SeedSequence(seed, spawn_key=(c,)))), so that the chunks can be filled by several host
    threads (numpy shuffles without the GIL; 1000 x 100 000 indices: 2.7 s on one core) and a seed still fixes every
    index.  seed=None draws fresh entropy once."""
    from concurrent.futures import ThreadPoolExecutor
    from .parallel import host_threads
    P, N = int(num_permutations), int(num_sequences)
    if out is None:
        out = np.empty((P, N), dtype=np.int32)
    entropy = np.random.SeedSequence(seed).entropy
    base = np.arange(N, dtype=np.int32)

    def fill(c):
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(entropy, spawn_key=(c,))))
        for p in range(c * PERM_CHUNK, min(P, (c + 1) * PERM_CHUNK)):
            out[p] = rng.permutation(base)

    n_chunks = -(-P // PERM_CHUNK)
    threads = host_threads() if threads is None else int(threads)
    if threads <= 1 or n_chunks <= 1 or P * N < (1 << 22):
        for c in range(n_chunks):
            fill(c)
    else:
        with ThreadPoolExecutor(max_workers=min(threads, n_chunks)) as ex:
            list(ex.map(fill, range(n_chunks)))
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

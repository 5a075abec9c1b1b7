"""Permuted-score columns (reference: mepp/permutations.py:5-58).

The reference builds an (N, 1+P) float32 matrix whose column 0 is the true score and whose
other columns are tf.random.shuffle(scores) -- unseeded (permutations.py:33-37).  Here the
permutations are an explicit (P, N) int32 index matrix generated on the host (optionally
seeded) so that the CPU oracle and the GPU path see identical permutations; the (N, 1+P)
matrix itself is only ever materialised on the GPU, as the tiled fp16 GEMM operand.
"""
import numpy as np


PERM_CHUNK = 64     # rows per independent random stream: the result does not depend on the number of threads


class PermutationSource:
    """The (P, N) int32 permutation index matrix as a generator of row ranges.

    Row p is a uniform random permutation; rows are drawn in chunks of PERM_CHUNK, chunk c from numpy
    Generator(PCG64(SeedSequence(seed, spawn_key=(c,)))), so any range of rows can be produced on its own -- by
    several host threads (numpy shuffles without the GIL; 1000 x 100 000 indices take 2.7 s on one core), and without
    ever holding the whole matrix: at cfg 5 (N = 1e6, P = 1e4) it is 40 GB, the engine stages it 256 rows at a time and
    builds the score operand Y behind it (Engine._yrows).  The reference builds the (N, 1 + P) matrix whole
    (mepp/permutations.py:27-39) and warns about its size (README.rst:53-56)."""

    def __init__(self, num_sequences, num_permutations, seed=None):
        self.N, self.P = int(num_sequences), int(num_permutations)
        self.entropy = np.random.SeedSequence(seed).entropy      # seed=None: fresh entropy, drawn once
        self.shape = (self.P, self.N)
        self.dtype = np.dtype(np.int32)

    def rows(self, p0, p1, out=None, threads=None):
        """Rows p0 .. p1-1 as a (p1 - p0, N) int32 array (p0 a multiple of PERM_CHUNK)."""
        from concurrent.futures import ThreadPoolExecutor
        from .parallel import host_threads
        p0, p1 = int(p0), min(int(p1), self.P)
        if p0 % PERM_CHUNK:
            raise ValueError(f"row ranges start at multiples of {PERM_CHUNK}")
        if out is None:
            out = np.empty((max(p1 - p0, 0), self.N), dtype=np.int32)
        base = np.arange(self.N, dtype=np.int32)

        def fill(c):
            rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(self.entropy, spawn_key=(c,))))
            for p in range(c * PERM_CHUNK, min(p1, (c + 1) * PERM_CHUNK)):
                out[p - p0] = rng.permutation(base)

        chunks = range(p0 // PERM_CHUNK, -(-p1 // PERM_CHUNK))
        threads = host_threads() if threads is None else int(threads)
        if threads <= 1 or len(chunks) <= 1 or (p1 - p0) * self.N < (1 << 22):
            for c in chunks:
                fill(c)
        else:
            with ThreadPoolExecutor(max_workers=min(threads, len(chunks))) as ex:
                list(ex.map(fill, chunks))
        return out

    def materialize(self, threads=None):
        return self.rows(0, self.P, threads=threads)


def make_permutation_indices(num_sequences, num_permutations, seed=None, threads=None, out=None):
    """(P, N) int32, row p a uniform random permutation: PermutationSource(N, P, seed) in full."""
    return PermutationSource(num_sequences, num_permutations, seed).rows(0, num_permutations, out=out, threads=threads)


LAZY_PERMUTATION_BYTES = 8 << 30     # index matrices above this size stay a PermutationSource (staged in chunks)


def add_permuted_scores_to_dataset(dataset, batch_size=None, num_permutations=1000, perm_indices=None, seed=None):
    """Same call shape as the reference (permutations.py:5-9); P <= 0 returns the dataset unchanged (:10-11)."""
    if perm_indices is None and num_permutations <= 0:
        return dataset
    if perm_indices is None:
        perm_indices = PermutationSource(dataset.num_sequences, num_permutations, seed)
        if 4 * perm_indices.P * perm_indices.N <= LAZY_PERMUTATION_BYTES:
            perm_indices = perm_indices.materialize()
    if not isinstance(perm_indices, PermutationSource):
        perm_indices = np.asarray(perm_indices, dtype=np.int32)
    if len(perm_indices.shape) != 2 or perm_indices.shape[1] != dataset.num_sequences:
        raise ValueError("perm_indices must be (P, N)")
    kw = dict(perm_idx=perm_indices)
    if batch_size is not None:
        kw['batch_size'] = batch_size
    return dataset.replace(**kw)

"""The staged dataset object that stands in for the reference's tf.data.Dataset.

In the reference a "dataset" is a batched tf.data pipeline of
(one-hot uint8 (B, L, 4), scores float32 (B,) or (B, 1+P), gc float32 (B, 1, 1))
(mepp/utils.py:182-194, mepp/permutations.py:42-56, mepp/utils.py:275-308).  Here it is a
plain host object -- ASCII sequences, float32 scores, an explicit permutation index matrix
and a "has GC covariate" flag -- plus a lazily created Engine that stages it once in HBM.
"""
import numpy as np


class ScoredDataset:
    def __init__(self, ascii_u8, scores, perm_idx=None, has_gc=False, batch_size=1000):
        self.ascii = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        if self.ascii.ndim != 2:
            raise ValueError("ascii must be (N, L) uint8")
        self.scores = np.ascontiguousarray(scores, dtype=np.float32)
        if self.scores.shape != (self.ascii.shape[0],):
            raise ValueError("scores must be (N,)")
        # (P, N) int32 indices, or a permutations.PermutationSource that generates them in row ranges
        self.perm_idx = perm_idx if perm_idx is None or hasattr(perm_idx, 'rows') else \
            np.ascontiguousarray(perm_idx, dtype=np.int32)
        self.has_gc = bool(has_gc)
        self.batch_size = batch_size
        self._engine = None

    # -- what the reference reads off element_spec (core.py:30-32, utils.py:196-200)
    @property
    def num_sequences(self):
        return self.ascii.shape[0]

    @property
    def sequence_length(self):
        return self.ascii.shape[1]

    @property
    def num_permutations(self):
        return 0 if self.perm_idx is None else self.perm_idx.shape[0]

    @property
    def element_spec(self):
        spec = [("onehot_sequences", (None, self.sequence_length, 4), "uint8")]
        if self.perm_idx is None:
            spec.append(("scores", (None,), "float32"))
        else:
            spec.append(("scores", (None, 1 + self.num_permutations), "float32"))
        if self.has_gc:
            spec.append(("gc_ratios", (None, 1, 1), "float32"))
        return tuple(spec)

    def replace(self, **kw):
        d = dict(ascii_u8=self.ascii, scores=self.scores, perm_idx=self.perm_idx, has_gc=self.has_gc,
                 batch_size=self.batch_size)
        d.update(kw)
        return ScoredDataset(**d)

    def engine(self, **kw):
        """The Engine holding this dataset staged in HBM (created on first use)."""
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(**kw).stage(self.ascii, self.scores, self.perm_idx, use_gc=self.has_gc)
        return self._engine

    def release(self):
        self._engine = None

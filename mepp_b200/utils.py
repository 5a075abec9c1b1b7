"""Staging helpers with the reference's names and semantics (reference: mepp/utils.py:32-308).

filter / order follow utils.py:77-128; the one-hot + tf.data construction of utils.py:141-194
becomes an ASCII matrix inside a ScoredDataset (the 2-bit pack happens on the GPU), and the GC
covariate of utils.py:275-308 becomes a flag: z is computed by the pack kernel at staging.
"""
import numpy as np
import pandas as pd

from .dataset import ScoredDataset
from .onehot_dna import get_degenerate_pct, seqs_to_ascii

# byte -> 1 for every character that is not an upper-case A, C, G, T (onehot_dna.py:41-53)
_NOT_ACGT_TABLE = bytes(0 if chr(b) in 'ACGT' else 1 for b in range(256))


def force_cpu_only():
    """utils.py:32-38 switched TensorFlow to the CPU.  This engine has no CPU path."""
    raise RuntimeError("mepp_b200 has no CPU fallback: --nogpu / no_gpu=True is not supported")


def manage_gpu_memory(device_index=None):
    """utils.py:41-51 enabled TF memory growth; nothing to do: the engine plans its memory up front."""
    return None


def scored_fasta_dicts_to_df(sequence_dict, score_dict, description_dict):
    """utils.py:53-75 (a sequence without a score raises KeyError, like the reference)."""
    ids = list(sequence_dict.keys())       # (column lists: pandas does not have to transpose 100 000 row tuples)
    return pd.DataFrame({'sequence_id': ids, 'sequence': list(sequence_dict.values()),
                         'score': [score_dict[sid] for sid in ids],
                         'description': [description_dict[sid] for sid in ids]},
                        columns=['sequence_id', 'sequence', 'score', 'description'])


def filter_scored_fasta_df(scored_fasta_df, degenerate_pct_thresh=100, sequence_length='max'):
    """utils.py:77-99."""
    df = scored_fasta_df.copy()
    df['length'] = df['sequence'].map(len)
    if sequence_length == 'max':
        sequence_length = df['length'].max()
    if sequence_length is not None:
        df = df[df['length'] == sequence_length].copy()
    seqs = list(df['sequence'])
    if sequence_length is not None and seqs:
        # all rows have one length now: the share of characters outside ACGT (upper case; onehot_dna.py:41-53) of every
        # sequence from one byte matrix instead of one Python call per sequence
        # (one bytes.translate over the joined sequences marks the characters outside ACGT; summing the flags of a row
        # counts them -- no index lookup per base)
        L_ = int(sequence_length)
        buf = ''.join(seqs).encode('latin-1', 'replace')
        if len(buf) == len(seqs) * L_:
            flags = np.frombuffer(buf.translate(_NOT_ACGT_TABLE), dtype=np.uint8).reshape(len(seqs), L_)
            df['degenerate_pct'] = 100.0 * flags.sum(axis=1, dtype=np.int64) / float(sequence_length)
        else:                                   # (rows of unequal byte length: per-sequence path)
            df['degenerate_pct'] = df['sequence'].map(get_degenerate_pct)
    else:
        df['degenerate_pct'] = df['sequence'].map(get_degenerate_pct)
    return df[df['degenerate_pct'] <= degenerate_pct_thresh].copy().reset_index(drop=True)


def order_scored_fasta_df(scored_fasta_df, shuffle=True, sort=True, ascending=True, sort_column='score',
                          random_state=None):
    """utils.py:101-128.  The reference's shuffle is unseeded; random_state makes it reproducible."""
    df = scored_fasta_df.copy()
    if shuffle:
        df = df.sample(frac=1, random_state=random_state).copy().reset_index(drop=True)
    if sort:
        df = df.sort_values(by=sort_column, ascending=ascending).copy().reset_index(drop=True)
        df['rank'] = np.argsort(df[sort_column])
    return df


def scored_fasta_df_to_dataset(scored_fasta_df, batch_size=1000, sequence_column='sequence',
                               score_column='score', n_jobs=1, progress_wrapper=None, **_ignored):
    """utils.py:141-194: rows -> dataset of (sequence, float32 score)."""
    seqs = list(scored_fasta_df[sequence_column])
    ascii_u8 = seqs_to_ascii(seqs)
    scores = np.asarray(scored_fasta_df[score_column].astype(float), dtype=np.float32)
    return ScoredDataset(ascii_u8, scores, batch_size=batch_size)


def get_input_shape_from_dataset(dataset):
    return (dataset.sequence_length, 4)


def get_sequence_length_from_dataset(dataset):
    return dataset.sequence_length


def append_gc_ratios_to_dataset(dataset, gc_ratio_type='global'):
    """utils.py:275-308.  Only the 'global' covariate (what run_mepp hard-codes, mepp.py:96) exists."""
    if gc_ratio_type != 'global':
        raise NotImplementedError("only gc_ratio_type='global' is supported (mepp/mepp.py:96)")
    return dataset.replace(has_gc=True)

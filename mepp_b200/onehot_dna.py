"""Alphabet rules of the reference's mepp/onehot_dna.py, kept for API compatibility.

The engine never materialises one-hot arrays: sequences travel as ASCII and are packed to
2-bit codes + an N mask on the GPU (csrc/pack.cu).  These helpers keep the reference's
semantics: channel order A,C,G,T (onehot_dna.py:3), every character that is not an
upper-case ACGT becomes N = [1,1,1,1] (onehot_dna.py:23-30), degenerate percentage counts
the same characters (onehot_dna.py:41-53).
"""
import numpy as np

nucs = list('ACGT')
alphabet = nucs.copy()
nuc_to_vec = {nuc: [1 if cur is nuc else 0 for cur in nucs] for nuc in nucs}
nuc_to_vec['N'] = [1, 1, 1, 1]
nuc_to_vec['_'] = [0, 0, 0, 0]
vec_to_nuc = {tuple(v): k for k, v in nuc_to_vec.items()}
nucs.append('N')
nucs.append('_')

_IS_ACGT = np.zeros(256, dtype=bool)
for _c in 'ACGT':
    _IS_ACGT[ord(_c)] = True


def masked_seq_to_upper(seq):
    return seq.upper()


def masked_seq_to_n_mask(seq, alphabet=alphabet):
    return ''.join(nuc if nuc in alphabet else 'N' for nuc in seq)


def masked_seq_to_zero_mask(seq, alphabet=alphabet):
    return ''.join(nuc if nuc in alphabet else '_' for nuc in seq)


def get_degenerate_pct(seq, alphabet=alphabet):
    """onehot_dna.py:41-53 (vectorised for the default alphabet)."""
    if not len(seq):
        return 0.0
    if alphabet is globals()['alphabet'] or list(alphabet) == list('ACGT'):
        b = np.frombuffer(seq.encode('latin-1', 'replace'), dtype=np.uint8)
        return 100.0 * float((~_IS_ACGT[b]).sum()) / len(seq)
    return 100.0 * sum(1 for nuc in seq if nuc not in alphabet) / len(seq)


convert_masked_seq_dict = {'upper': masked_seq_to_upper, 'n': masked_seq_to_n_mask, '_': masked_seq_to_zero_mask}


def seq_to_mat(seq, nuc_to_vec=nuc_to_vec, convert_masked_seq=convert_masked_seq_dict['n']):
    return [nuc_to_vec[nuc] for nuc in convert_masked_seq(seq)]


def mat_to_seq(mat, vec_to_nuc=vec_to_nuc):
    return ''.join(vec_to_nuc[tuple(vec)] for vec in list(mat))


def seqs_to_ascii(seqs, length=None):
    """List of equal-length strings -> (N, L) uint8 ASCII matrix (the engine's input format)."""
    seqs = list(seqs)
    if length is None:
        length = len(seqs[0]) if seqs else 0
    buf = ''.join(seqs).encode('latin-1', 'replace')
    if len(buf) != len(seqs) * length:
        raise ValueError("all sequences must have the same length")
    return np.frombuffer(buf, dtype=np.uint8).reshape(len(seqs), length).copy()

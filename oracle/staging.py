"""Oracle restatement of the reference's staging (L0) layer.  Test infrastructure only.

Follows mepp/io.py:19-95 (JASPAR-style motif parse through Biopython, scored
FASTA parse), mepp/onehot_dna.py:3-70 (one-hot, N rule, degenerate pct),
mepp/utils.py:53-128 (filter / order), mepp/utils.py:202-308 (GC covariate),
mepp/permutations.py:27-39 (permuted score matrix).
"""
import re
import numpy as np

NUCS = "ACGT"  # onehot_dna.py:3 channel order


def parse_motifs(text):
    """mepp/io.py:19-46 + Bio.motifs.jaspar.read(format='jaspar') semantics.

    Header regex ``^>\\s*(\\S+)(\\s+(\\S+))?`` -> (matrix_id, name or matrix_id);
    unlabelled rows are taken in order A,C,G,T; ``motif.pwm`` is counts divided by
    the column sum; dict key is ``"{id} {name}"`` unless id == name (io.py:29);
    later duplicates overwrite earlier ones (dict semantics).
    Returns {key: (4, m) float64}.
    """
    out = {}
    key, rows = None, []

    def flush():
        if key is not None:
            counts = np.array(rows, dtype=np.float64)
            assert counts.shape[0] == 4, f"motif {key!r}: expected 4 rows"
            out[key] = counts / counts.sum(axis=0, keepdims=True)

    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            flush()
            m = re.match(r"^>\s*(\S+)(\s+(\S+))?", line)
            mid = m.group(1)
            name = m.group(3) if m.group(3) else mid
            key = f"{mid} {name}" if mid != name else mid
            rows = []
        else:
            # JASPAR rows may be labelled "A [ 1 2 3 ]"; the shipped files are bare numbers
            toks = re.sub(r"[\[\]]", " ", line).split()
            if toks and toks[0] in ("A", "C", "G", "T"):
                toks = toks[1:]
            rows.append([float(t) for t in toks])
    flush()
    return out


def parse_scored_fasta(text, description_delim=" "):
    """mepp/io.py:58-84: id = first token, score = float(description.split(' ')[1]).

    Returns (sequence_dict, score_dict, description_dict) with dict (last wins)
    semantics, like the reference.
    """
    seqs, scores, descs = {}, {}, {}
    rec_id, desc, chunks = None, None, []

    def flush():
        if rec_id is not None:
            seqs[rec_id] = "".join(chunks)
            parts = desc.split(description_delim)
            if len(parts) > 1:
                descs[rec_id] = tuple(parts)
                scores[rec_id] = float(parts[1])

    for line in text.splitlines():
        if line.startswith(">"):
            flush()
            desc = line[1:].strip()
            rec_id = desc.split()[0] if desc.split() else ""
            chunks = []
        else:
            chunks.append(line.strip())
    flush()
    return seqs, scores, descs


def degenerate_pct(seq):
    """onehot_dna.py:41-53: percentage of characters not in uppercase ACGT."""
    return 100.0 * sum(1 for c in seq if c not in NUCS) / len(seq)


def filter_and_order(seqs, scores, descs, degenerate_pct_thresh=100, shuffle_rng=None):
    """utils.py:53-128 as rows (sequence_id, sequence, score).

    Keeps max-length sequences, drops degenerate_pct > threshold, optionally
    shuffles (the reference's shuffle is unseeded; pass a Generator to mimic it),
    then sorts ascending by score (stable here; pandas' default is unstable).
    """
    rows = [(k, s, scores[k]) for k, s in seqs.items()]
    if not rows:
        return rows
    max_len = max(len(r[1]) for r in rows)
    rows = [r for r in rows if len(r[1]) == max_len]
    rows = [r for r in rows if degenerate_pct(r[1]) <= degenerate_pct_thresh]
    if shuffle_rng is not None:
        perm = shuffle_rng.permutation(len(rows))
        rows = [rows[i] for i in perm]
    rows.sort(key=lambda r: r[2])
    return rows


def onehot(seq):
    """onehot_dna.py:61-70 with the 'n' mask: non-ACGT(upper) -> [1,1,1,1]."""
    out = np.zeros((len(seq), 4), dtype=np.uint8)
    for i, c in enumerate(seq):
        k = NUCS.find(c)
        if k >= 0:
            out[i, k] = 1
        else:
            out[i, :] = 1
    return out


def ascii_to_codes(ascii_arr):
    """(N, L) uint8 ASCII -> int8 codes 0..3 for ACGT (uppercase only), 4 for anything else."""
    lut = np.full(256, 4, dtype=np.int8)
    for k, c in enumerate(NUCS):
        lut[ord(c)] = k
    return lut[ascii_arr]


def gc_ratio_global(codes):
    """utils.py:202-273: z = (1/L) sum_l (xhat[C] + xhat[G]); N contributes 0.5.  float32.

    The per-position values are 0, 0.5 or 1 so the float32 sum is exact; the
    division by L (a float32) rounds once.
    """
    per = np.where((codes == 1) | (codes == 2), 1.0, np.where(codes == 4, 0.5, 0.0)).astype(np.float32)
    s = per.sum(axis=1, dtype=np.float64).astype(np.float32)
    return s / np.float32(codes.shape[1])


def permuted_scores(scores_f32, perm_idx):
    """permutations.py:27-39: column 0 true scores, column p = scores[perm_idx[p-1]].

    perm_idx is an explicit (P, N) index matrix because the reference's
    tf.random.shuffle is unseeded.
    """
    cols = [scores_f32] + [scores_f32[perm_idx[p]] for p in range(perm_idx.shape[0])]
    return np.stack(cols, axis=1).astype(np.float32)

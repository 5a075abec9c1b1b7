"""Input parsing and dataset persistence (reference: mepp/io.py).

* motif_matrix_file_to_dicts  -- mepp/io.py:19-46, restating Bio.motifs.jaspar.read(format='jaspar')
  for the plain ">id name" + four unlabelled (or "A [ ... ]" labelled) count rows the MEPP motif
  libraries use: id/name from ``^>\\s*(\\S+)(\\s+(\\S+))?``, rows in order A,C,G,T, pwm = counts /
  column sum, key "{id} {name}" unless they are equal, later duplicates overwrite earlier ones.
* scored_fasta_file_to_dicts  -- mepp/io.py:58-84: id = first token of the header, score =
  float(description.split(' ')[1]).
* save_dataset / load_dataset -- mepp/io.py:97-133 with the same call shape; the TF snapshot
  container is replaced by one .npz inside the dataset directory, and the pickled
  ``.element_spec.p`` sidecar is kept for layout parity.
"""
import os
import pickle as p
import re
import shutil

import numpy as np

from .dataset import ScoredDataset
from .onehot_dna import alphabet

_HEADER = re.compile(r"^>\s*(\S+)(\s+(\S+))?")


def _consensus(pwm):
    return ''.join('ACGT'[i] for i in np.argmax(pwm, axis=0))


def motif_matrix_file_to_dicts(motifs_file, alphabet=alphabet):
    motif_matrix_dict, motif_consensus_dict = {}, {}
    state = dict(mid=None, name=None, rows=[])

    def flush():
        if state['mid'] is None:
            return
        counts = np.array(state['rows'], dtype=np.float64)
        if counts.ndim != 2 or counts.shape[0] != 4:
            raise ValueError(f"motif {state['mid']!r}: expected 4 rows of equal length")
        pwm = counts / counts.sum(axis=0, keepdims=True)
        mid, name = state['mid'], state['name']
        key = f'{mid} {name}' if mid != name else mid
        motif_matrix_dict[key] = np.array([list(pwm['ACGT'.index(nuc)]) for nuc in alphabet])
        motif_consensus_dict[f'{mid} {name}'] = _consensus(pwm)

    for line in motifs_file:
        line = line.strip()
        if not line:
            continue
        if line.startswith('>'):
            flush()
            m = _HEADER.match(line)
            state['mid'] = m.group(1)
            state['name'] = m.group(3) if m.group(3) else m.group(1)
            state['rows'] = []
        else:
            toks = re.sub(r"[\[\]]", " ", line).split()
            if toks and toks[0] in ('A', 'C', 'G', 'T'):
                toks = toks[1:]
            state['rows'].append([float(t) for t in toks])
    flush()
    return motif_matrix_dict, motif_consensus_dict


def motif_matrix_filepath_to_dicts(motifs_filepath, **kwargs):
    with open(motifs_filepath) as motifs_file:
        return motif_matrix_file_to_dicts(motifs_file, **kwargs)


def scored_fasta_file_to_dicts(fasta_file, description_delim=' '):
    sequence_dict, description_dict, score_dict = {}, {}, {}
    rec = dict(id=None, desc=None, chunks=[])

    def flush():
        if rec['id'] is None:
            return
        sequence_dict[rec['id']] = ''.join(rec['chunks'])
        parts = rec['desc'].split(description_delim)
        if len(parts) > 1:
            description_dict[rec['id']] = tuple(parts)
            score_dict[rec['id']] = float(parts[1])

    for line in fasta_file:
        if line.startswith('>'):
            flush()
            rec['desc'] = line[1:].strip()
            toks = rec['desc'].split()
            rec['id'] = toks[0] if toks else ''
            rec['chunks'] = []
        else:
            rec['chunks'].append(line.strip())
    flush()
    return sequence_dict, score_dict, description_dict


def scored_fasta_filepath_to_dicts(fasta_filepath, **kwargs):
    import click
    with click.open_file(fasta_filepath) as fasta_file:   # '-' reads stdin (io.py:91)
        return scored_fasta_file_to_dicts(fasta_file, **kwargs)


def save_dataset(dataset, path, compression=None, element_spec_suffix='.element_spec.p', delete_existing=False,
                 arrays=True):
    """io.py:97-115.  arrays=False writes only the element_spec sidecar (run_batch, when the dataset directory is
    removed again right after the run)."""
    normpath = os.path.normpath(path)
    if delete_existing and os.path.exists(normpath):
        shutil.rmtree(normpath)
    os.makedirs(normpath, exist_ok=True)
    if not arrays:
        with open(os.path.normpath(f'{path}{element_spec_suffix}'), 'wb') as f:
            p.dump(dataset.element_spec, f)
        return
    arrays = dict(ascii=dataset.ascii, scores=dataset.scores, has_gc=np.array(dataset.has_gc),
                  batch_size=np.array(dataset.batch_size if dataset.batch_size is not None else -1))
    if dataset.perm_idx is not None and hasattr(dataset.perm_idx, 'rows'):
        # a PermutationSource is saved as its seed entropy: the indices are regenerated in chunks when staged
        arrays['perm_source'] = np.array([str(dataset.perm_idx.entropy), str(dataset.perm_idx.P)])
    elif dataset.perm_idx is not None:
        arrays['perm_idx'] = dataset.perm_idx
    saver = np.savez_compressed if compression else np.savez
    saver(os.path.join(normpath, 'dataset.npz'), **arrays)
    with open(os.path.normpath(f'{path}{element_spec_suffix}'), 'wb') as f:
        p.dump(dataset.element_spec, f)


def load_dataset(path, compression=None, element_spec_suffix='.element_spec.p'):
    with np.load(os.path.join(os.path.normpath(path), 'dataset.npz')) as z:
        bs = int(z['batch_size'])
        perm = z['perm_idx'] if 'perm_idx' in z else None
        if 'perm_source' in z:
            from .permutations import PermutationSource
            perm = PermutationSource(z['ascii'].shape[0], int(z['perm_source'][1]), seed=int(z['perm_source'][0]))
        return ScoredDataset(z['ascii'], z['scores'], perm,
                             has_gc=bool(z['has_gc']), batch_size=None if bs < 0 else bs)

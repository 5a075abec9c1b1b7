"""Synthetic scored-FASTA workloads of the shapes named in BASELINE.json.

numpy Generator(PCG64(seed)); iid uniform ACGT with 0.2 % N and 0.2 % lower-cased bases,
scores ~ N(5, 2) rounded to 6 decimals, rows pre-sorted ascending by score (stable), the
consensus of the first motifs planted near the centre with score-dependent probability so
that profiles are not null; permutations from PCG64(seed + 1).
"""
import numpy as np

CONFIGS = {
    # name: (N, L, motif file, orientations, P, margin)
    "cfg1": (1000, 201, "ohler_motifs.txt", ['+', '+/-'], 100, 2),
    "cfg2": (20000, 1000, "homer.motifs.clustered.164_nonredundant.txt", ['+', '+/-'], 1000, 2),
    "cfg3": (100000, 1000, "homer.motifs.clustered.164_nonredundant.txt", ['+', '+/-'], 1000, 2),
    "cfg4": (200000, 500, "homer.motifs.txt", ['+', '-', '+/-'], 1000, 2),
    "cfg5": (1000000, 200, "homer.motifs.clustered.164_nonredundant.txt", ['+', '+/-'], 10000, 2),
}
SEED_BASE = 20240


def synth_sequences(N, L, seed, motifs=None, n_plant=8, n_rate=0.002, lower_rate=0.002):
    """Returns (ascii (N, L) uint8, scores (N,) float32 sorted ascending)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scores = np.round(rng.normal(5.0, 2.0, size=N), 6)
    scores = np.sort(scores, kind='stable').astype(np.float32)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = letters[rng.integers(0, 4, size=(N, L), dtype=np.uint8)]
    if motifs:
        center = L // 2
        for pwm in list(motifs)[:n_plant]:
            m = pwm.shape[1]
            if m + 24 > L:
                continue
            cons = letters[np.argmax(pwm, axis=0)]
            prob = 0.3 / (1.0 + np.exp(-(scores.astype(np.float64) - 5.0)))
            who = np.nonzero(rng.random(N) < prob)[0]
            starts = center + rng.integers(-10, 11, size=who.size) - m // 2
            starts = np.clip(starts, 0, L - m)
            for j in range(m):
                seq[who, starts + j] = cons[j]
    r = rng.random((N, L))
    seq = np.where(r < n_rate, np.uint8(ord('N')), seq)
    lower = (r >= n_rate) & (r < n_rate + lower_rate)
    seq = np.where(lower, seq | np.uint8(0x20), seq).astype(np.uint8)
    return np.ascontiguousarray(seq), scores


LAZY_PERM_BYTES = 8 << 30     # above this the index matrix stays a PermutationSource (cfg 5: 40 GB)


def synth_permutations(N, P, seed, lazy=None):
    """(P, N) int32 permutation indices from PCG64 streams seeded with seed + 1 (permutations.PermutationSource: chunked
    streams, filled by the host threads).  Matrices above LAZY_PERM_BYTES are returned as the source itself: the engine
    stages it in row chunks and the oracle asks for the rows it checks."""
    from .permutations import PermutationSource
    src = PermutationSource(N, P, seed + 1)
    if lazy is None:
        lazy = 4 * N * P > LAZY_PERM_BYTES
    return src if lazy else src.materialize()


def synth_motifs(n_motifs, seed, m_lo=6, m_hi=20):
    """Random PWMs (Dirichlet columns) for tests that must not read the reference's data files."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(n_motifs):
        m = int(rng.integers(m_lo, m_hi + 1))
        cols = rng.dirichlet(np.full(4, 0.3), size=m).T
        out.append(np.ascontiguousarray(cols, dtype=np.float64))
    return out


def load_motif_library(name, fixtures=None):
    """Motif library from mepp_b200/data/motif_libraries.npz: the libraries the reference ships in its data/ directory,
    packaged with this path (made from those files by tests/golden/make_motif_fixtures.py).
    name in {'ohler', 'homer_clustered', 'homer'}.
    Returns {key: (4, m) float64} with the reference's rule that a later duplicate key overwrites an
    earlier one (mepp/io.py:27-38) -- 164 headers -> 162 keys for 'homer_clustered'."""
    import os
    if fixtures is None:
        fixtures = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "motif_libraries.npz")
    with np.load(fixtures) as z:
        keys, lens, flat = z[f"{name}__keys"], z[f"{name}__lens"], z[f"{name}__flat"]
    out, off = {}, 0
    for k, m in zip(keys, lens):
        out[str(k)] = flat[off:off + 4 * m].reshape(4, m).copy()
        off += 4 * m
    return out


def make_config(cfg, n_seq=None, n_motifs=None, n_perm=None):
    """Synthetic workload of one BASELINE.json config (optionally shrunk for tests).
    Returns dict(ascii, scores, perm_idx, tasks=[(key, pwm, orientation)], margin, L, N, P)."""
    N, L, lib, orients, P, margin = CONFIGS[cfg]
    idx = int(cfg[3:])
    seed = SEED_BASE + idx
    libname = {"ohler_motifs.txt": "ohler", "homer.motifs.clustered.164_nonredundant.txt": "homer_clustered",
               "homer.motifs.txt": "homer"}[lib]
    motifs = load_motif_library(libname)
    if n_seq is not None:
        N = n_seq
    if n_perm is not None:
        P = n_perm
    keys = list(motifs.keys())
    if n_motifs is not None:
        keys = keys[:n_motifs]
    ascii_u8, scores = synth_sequences(N, L, seed, motifs=[motifs[k] for k in keys])
    perm_idx = synth_permutations(N, P, seed)
    tasks = [(k, motifs[k], o) for k in keys for o in orients]   # batch.py:73-90 order: motif outer, orientation inner
    return dict(ascii=ascii_u8, scores=scores, perm_idx=perm_idx, tasks=tasks, margin=margin, L=L, N=N, P=P,
                cfg=cfg)

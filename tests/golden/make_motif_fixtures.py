"""Generates mepp_b200/data/motif_libraries.npz (package data: bench / tests / synthetic workloads read the shipped
libraries from there, not from tests/) from the reference's motif libraries.

Run in the build container (needs /root/reference); the GPU box only has the committed
.npz.  Stores, per library, the headers in FILE ORDER (duplicates kept, so the reference's
"later duplicate key wins" rule can be exercised) and the column-normalised 4 x m PWMs.
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import staging  # noqa: E402

REF = "/root/reference/data"
LIBS = {"ohler": "ohler_motifs.txt", "homer_clustered": "homer.motifs.clustered.164_nonredundant.txt",
        "homer": "homer.motifs.txt"}


def parse_in_order(text):
    keys, mats = [], []
    cur = None
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            cur = []
            d = staging.parse_motifs(line + "\n1 1\n1 1\n1 1\n1 1\n")
            keys.append(list(d.keys())[0])
            mats.append(cur)
        else:
            cur.append([float(t) for t in line.split()])
    pw = []
    for m in mats:
        c = np.array(m, dtype=np.float64)
        pw.append(c / c.sum(axis=0, keepdims=True))
    return keys, pw


out = {}
for lib, fn in LIBS.items():
    keys, pw = parse_in_order(open(os.path.join(REF, fn)).read())
    out[f"{lib}__keys"] = np.array(keys)
    out[f"{lib}__lens"] = np.array([p.shape[1] for p in pw], dtype=np.int32)
    out[f"{lib}__flat"] = np.concatenate([p.ravel() for p in pw])
    print(lib, len(keys), "headers,", len(set(keys)), "unique keys")
np.savez_compressed(os.path.join(os.path.dirname(__file__), "..", "..", "mepp_b200", "data", "motif_libraries.npz"), **out)

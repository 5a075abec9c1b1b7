"""Lean scan (Engine.lean, mepp_scan_lean) against the full scan: identical matches / counts / densities, r within
1e-6, bucket overflow falls back to the full scan; then the cfg 2 step time with and without it (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

ok = True
for case, (N, L, margin, seed) in enumerate([(333, 77, 0, 1), (333, 150, 2, 2), (1000, 201, 5, 3), (257, 64, 16, 4)]):
    ascii_u8, scores = synth.synth_sequences(N, L, seed=seed, n_rate=0.02, lower_rate=0.01)
    if case == 0:
        ascii_u8[6, :] = ord('A')                                        # a long run of matches: bucket overflow -> full scan
    else:
        ascii_u8[6, 20:32] = ord('A')                                    # a short run of overlapping matches
    perm = synth.synth_permutations(N, 13, seed=seed)
    rng = np.random.default_rng(seed)
    motifs = synth.synth_motifs(6, seed=seed, m_lo=4, m_hi=min(32, L // 3))
    cons = np.full((4, 9), 1e-3); cons[0, :] = 1.0
    motifs.append(cons / cons.sum(axis=0))
    tasks = [(m, o) for m in motifs for o in ('+', '+/-')] + [(motifs[0], '-')]
    res = []
    for lean in (False, True):
        eng = Engine()
        eng.lean = lean
        eng.stage(ascii_u8, scores, perm)
        res.append(eng.profile(tasks, margin=margin, return_hits=True, return_r=True))
        res[-1].paths = list(eng.last_paths)
    a, b = res
    same = (np.array_equal(a.hits, b.hits) and np.array_equal(a.density, b.density)
            and np.array_equal(np.asarray(a.positions['count']), np.asarray(b.positions['count'])))
    dr = float(np.abs(a.r - b.r).max())
    tol = dict(rtol=1e-6, atol=1e-9) if a.paths == b.paths else dict(rtol=1e-4, atol=1e-6)   # (fallback = tensor path)
    close = np.allclose(a.r, b.r, **tol)
    print(f"case {case} (N={N} L={L} margin={margin}): matches {len(a.hits)} equal {same}; max |dr| {dr:.2e} close {close}; "
          f"paths {b.paths}", flush=True)
    ok &= same and close
print("LEAN PARITY", "OK" if ok else "FAILED", flush=True)

cfg = synth.make_config("cfg2")
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
ref = None
for lean in (False, True, False, True):
    eng = Engine()
    eng.lean = lean
    eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    tables = eng.upload_tables(motif_layers.scan_tables(tasks))
    for _ in range(3):
        b = eng.profile(tasks, margin=cfg["margin"], tables=tables)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        b = eng.profile(tasks, margin=cfg["margin"], tables=tables)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / 10
    eng.overlap = False
    eng.timing = {}
    eng.profile(tasks, margin=cfg["margin"], tables=tables)
    torch.cuda.synchronize()
    tm = {k: round(v, 2) for k, v in eng.collect_timing().items()}
    sig = (b.density.copy(), np.asarray(b.positions['count']).copy())
    r0 = np.asarray(b.positions['positional_r']).copy()
    extra = ""
    if ref is None:
        ref = (sig, r0)
    else:
        extra = f" counts/densities equal {all(np.array_equal(x, y) for x, y in zip(sig, ref[0]))}, max |dr0| {np.abs(r0 - ref[1]).max():.2e}"
    print(f"lean {lean}: {ms:.2f} ms per step; serial stages {tm}; paths {set(eng.last_paths)}{extra}", flush=True)
    del eng, tables
    torch.cuda.empty_cache()

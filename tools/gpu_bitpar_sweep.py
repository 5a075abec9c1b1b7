"""Scan time of a whole config as a function of the cost ratio that admits units to the bit-parallel first stage, and
equality of the outputs with the pair-table-only scan (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config(os.environ.get("SCALE_CFG", "cfg2"), n_motifs=int(os.environ.get("SCALE_MOTIFS", "0")) or None)
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
cuts = [float(c) for c in os.environ.get("BP_RATIOS", "0,0.25,0.5,0.75,1,1.5,2,1e9").split(",")]
ref = None
for cut in cuts:
    eng = Engine()
    eng.bitpar = True
    eng.overlap = False          # clean per-stage times
    eng.bp_cost_ratio = cut
    eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    tables = eng.upload_tables(motif_layers.scan_tables(tasks))
    bp = tables.dev['bp'].cpu().numpy().view(np.uint32)
    use = (bp[:, 64] >> 8) & 1
    rate = bp[:, 66].copy().view(np.float32)
    taps = (bp[:, 65] & 0xff) + ((bp[:, 65] >> 8) & 0xff)
    best = None
    for rep in range(4):
        eng.timing = {}
        b = eng.profile(tasks, margin=cfg["margin"], tables=tables)
        torch.cuda.synchronize()
        tm = eng.collect_timing()
        if rep and (best is None or tm['scan'] < best['scan']):
            best = tm
    sig = (b.hit_counts.copy(), b.density.copy(), np.asarray(b.positions['count']).copy(),
           np.asarray(b.positions['positional_r']).copy())
    same = ""
    if ref is None:
        ref = sig
    else:
        same = " equal to first: %s" % all(np.array_equal(x, y) for x, y in zip(sig, ref))
    print(f"ratio {cut:g}: bit-parallel tasks {int(use.sum())}/{len(tasks)}  K hist {np.bincount(bp[:, 64] & 0xff, minlength=4).tolist()} "
          f"median rate {np.median(rate):.4f} mean counted taps {taps.mean():.1f}  stages(ms) { {k: round(v, 2) for k, v in best.items()} }{same}", flush=True)
    del eng, tables
    torch.cuda.empty_cache()

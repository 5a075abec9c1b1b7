"""Device-resident step with and without the scan / statistics overlap (Engine.overlap): time per step and equality
of every output (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config(os.environ.get("SCALE_CFG", "cfg2"), n_motifs=int(os.environ.get("SCALE_MOTIFS", "0")) or None)
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
eng = Engine()
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
tables = eng.upload_tables(motif_layers.scan_tables(tasks))
ref = None
for mode in os.environ.get("OVERLAP_MODES", "0,1,0,1").split(","):
    eng.overlap = mode == "1"
    for _ in range(3):
        b = eng.profile(tasks, margin=cfg["margin"], tables=tables)
    torch.cuda.synchronize()
    n = int(os.environ.get("OVERLAP_STEPS", "10"))
    t0 = time.perf_counter()
    for _ in range(n):
        b = eng.profile(tasks, margin=cfg["margin"], tables=tables)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / n
    sig = [b.hit_counts.copy(), b.density.copy()] + [np.asarray(v).copy() for v in b.positions.values()]
    same = ""
    if ref is None:
        ref = sig
    else:
        same = " outputs equal to the first mode: %s" % all(np.array_equal(x, y, equal_nan=True) for x, y in zip(sig, ref))
    print(f"overlap {mode}: {ms:.2f} ms per step ({cfg['N'] * cfg['L'] * len(tasks) / ms / 1e6:.0f} Gbp*motifs/s){same}", flush=True)

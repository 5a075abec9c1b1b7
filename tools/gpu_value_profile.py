"""Host-side breakdown of one device-resident step (run under gpurun)."""
import os, sys, time, cProfile, pstats, io
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config(os.environ.get("SCALE_CFG", "cfg2"))
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
eng = Engine()
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
tables = motif_layers.scan_tables(tasks)
for _ in range(3):
    eng.profile(tasks, margin=cfg["margin"], tables=tables)
torch.cuda.synchronize()
for rep in range(3):
    eng.timing = {}
    t0 = time.perf_counter(); eng.profile(tasks, margin=cfg["margin"], tables=tables); torch.cuda.synchronize()
    t1 = time.perf_counter()
    tm = eng.collect_timing()
    print(f"step {1e3*(t1-t0):.1f} ms; gpu stages {sum(tm.values()):.1f} ms {tm}", flush=True)
eng.timing = None
pr = cProfile.Profile(); pr.enable(); eng.profile(tasks, margin=cfg["margin"], tables=tables); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue())

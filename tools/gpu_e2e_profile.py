"""Host-side breakdown of one end-to-end step (run under gpurun)."""
import os, sys, time, cProfile, pstats, io
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config(os.environ.get("SCALE_CFG", "cfg2"))
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
h_ascii = torch.from_numpy(cfg["ascii"]).pin_memory()
h_scores = torch.from_numpy(cfg["scores"]).pin_memory()
h_perm = torch.from_numpy(cfg["perm_idx"]).pin_memory()
eng = Engine()


def step():
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
    tables = motif_layers.scan_tables(tasks)
    return eng.profile(tasks, margin=cfg["margin"], tables=tables)


for _ in range(3):
    step()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); eng.stage(h_ascii, h_scores, h_perm, use_gc=True); torch.cuda.synchronize()
    t1 = time.perf_counter(); tables = motif_layers.scan_tables(tasks)
    t2 = time.perf_counter(); b = eng.profile(tasks, margin=cfg["margin"], tables=tables); torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"stage {1e3*(t1-t0):.1f} ms  tables {1e3*(t2-t1):.1f} ms  profile {1e3*(t3-t2):.1f} ms", flush=True)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(35); print(s.getvalue())

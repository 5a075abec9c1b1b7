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
    return eng.profile(tasks, margin=cfg["margin"])


for _ in range(3):
    step()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); eng.stage(h_ascii, h_scores, h_perm, use_gc=True); torch.cuda.synchronize()
    t1 = time.perf_counter()
    t2 = time.perf_counter(); b = eng.profile(tasks, margin=cfg["margin"]); torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"stage {1e3*(t1-t0):.1f} ms  tables {1e3*(t2-t1):.1f} ms  profile {1e3*(t3-t2):.1f} ms", flush=True)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue())

# back-to-back steps like bench.py's e2e region
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    print(f"3 back-to-back e2e steps: {1e3*(time.perf_counter()-t0)/3:.1f} ms/step", flush=True)
import threading
ts = []
orig_stage, orig_profile = eng.stage, eng.profile
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); ts.append((name, 1e3*(time.perf_counter()-t))); return r
    return w
eng.stage = timed("stage", orig_stage); eng.profile = timed("profile", orig_profile)
eng._launch_group = timed("launch", eng._launch_group); eng._collect_group = timed("collect", eng._collect_group)
import mepp_b200.motif_layers as ml
ml_scan = ml.scan_tables
ml.scan_tables = timed("tables", ml_scan)
for _ in range(2):
    step()
torch.cuda.synchronize()
print([(n, round(t, 1)) for n, t in ts], flush=True)

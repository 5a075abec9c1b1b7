"""Find which host call of an end-to-end step stalls when a step is slow (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine
import mepp_b200.engine as E

cfg = synth.make_config("cfg2")
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
h_ascii = torch.from_numpy(cfg["ascii"]).pin_memory()
h_scores = torch.from_numpy(cfg["scores"]).pin_memory()
h_perm = torch.from_numpy(cfg["perm_idx"]).pin_memory()
eng = Engine()
log = []
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); log.append((name, 1e3 * (time.perf_counter() - t))); return r
    setattr(obj, name, w)
for n in ("stage", "device_tables_for", "_launch_group", "_collect_group", "_ensure_y", "_yrows", "_colconst"):
    wrap(eng, n)
wrap(motif_layers, "prepare_tables")
import gc
for _ in range(5):
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True); eng.profile(tasks, margin=cfg["margin"])
torch.cuda.synchronize(); gc.collect(); gc.disable()
slow = 0
for step in range(int(os.environ.get("SPIKE_STEPS", "300"))):
    log.clear()
    t = time.perf_counter()
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
    t1 = time.perf_counter()
    b = eng.profile(tasks, margin=cfg["margin"])
    dt = 1e3 * (time.perf_counter() - t)
    if dt > 26:
        slow += 1
        print(f"step {step}: {dt:.1f} ms (stage call {1e3*(t1-t):.1f}):", [(n, round(v, 1)) for n, v in log if v > 1.5], flush=True)
print("slow steps:", slow)

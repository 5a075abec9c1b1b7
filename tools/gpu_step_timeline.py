"""Where the wall time of one device-resident step goes (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config("cfg2")
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
eng = Engine()
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
tables = motif_layers.scan_tables(tasks)
for _ in range(3):
    eng.profile(tasks, margin=cfg["margin"], tables=tables)
torch.cuda.synchronize()
marks = []
orig_launch, orig_collect = eng._launch_group, eng._collect_group
def launch(*a, **k):
    t = time.perf_counter(); r = orig_launch(*a, **k); marks.append(("launch", t, time.perf_counter())); return r
def collect(p, *a, **k):
    t = time.perf_counter(); p['done'].synchronize(); t1 = time.perf_counter()
    r = orig_collect(p, *a, **k); marks.append(("collect", t, time.perf_counter(), t1)); return r
eng._launch_group, eng._collect_group = launch, collect
for rep in range(3):
    marks.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    eng.profile(tasks, margin=cfg["margin"], tables=tables)
    e1.record(); t_ret = time.perf_counter(); torch.cuda.synchronize(); t_end = time.perf_counter()
    print(f"wall to return {1e3*(t_ret-t0):.2f} ms, to sync {1e3*(t_end-t0):.2f} ms, GPU first->last event {e0.elapsed_time(e1):.2f} ms")
    for m in marks:
        if m[0] == "launch":
            print(f"   launch  {1e3*(m[1]-t0):6.2f} -> {1e3*(m[2]-t0):6.2f}")
        else:
            print(f"   collect {1e3*(m[1]-t0):6.2f} -> {1e3*(m[2]-t0):6.2f}  (results on host at {1e3*(m[3]-t0):6.2f})")

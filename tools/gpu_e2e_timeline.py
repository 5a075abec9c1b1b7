"""Device timeline of ONE end-to-end step (cfg3): when the uploads, the scans and the profile chains of the task groups
start and end, relative to the start of the step (CUDA events; run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth
from mepp_b200.engine import Engine

cfg = synth.make_config(os.environ.get("TL_CFG", "cfg3"))
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
h_ascii = torch.from_numpy(cfg["ascii"]).pin_memory()
h_scores = torch.from_numpy(cfg["scores"]).pin_memory()
h_perm = torch.from_numpy(cfg["perm_idx"]).pin_memory()
eng = Engine()
for _ in range(3):
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
    eng.profile(tasks, margin=cfg["margin"])
torch.cuda.synchronize()
cur = torch.cuda.current_stream()
for rep in range(2):
    E = lambda: torch.cuda.Event(enable_timing=True)
    ev0 = E(); t0 = time.perf_counter(); ev0.record(cur)
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
    t1 = time.perf_counter()
    ev_stage = E(); ev_stage.record(cur)
    ev_perm = E(); ev_perm.record(eng._copy_stream)
    eng.timing = {}
    batch = eng.profile(tasks, margin=cfg["margin"])
    t2 = time.perf_counter()
    ev_end = E(); ev_end.record(cur)
    torch.cuda.synchronize()
    print(f"rep {rep}: host stage {1e3 * (t1 - t0):.2f} ms, host profile {1e3 * (t2 - t1):.2f} ms; device: staging kernels done "
          f"{ev0.elapsed_time(ev_stage):.2f}, permutation upload done {ev0.elapsed_time(ev_perm):.2f}, end {ev0.elapsed_time(ev_end):.2f}")
    for name, a, b in eng.timing.get('_pairs', []):
        print(f"   {name:12s} {ev0.elapsed_time(a):8.2f} -> {ev0.elapsed_time(b):8.2f}")
    eng.timing = None

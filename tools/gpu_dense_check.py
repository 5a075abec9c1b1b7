"""A dense score matrix (--motifpvalue 1e-2, --margin 16, honoured): Engine.choose_path takes the tcgen05 GEMM, which has to
beat the gather paths there (run under gpurun).  cfg2-sized dataset, 16 tasks."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth
from mepp_b200.engine import Engine

cfg = synth.make_config("cfg2", n_motifs=8)
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
kw = dict(margin=int(os.environ.get("DENSE_MARGIN", "16")), honour_flags=True, pvalue=float(os.environ.get("DENSE_PVAL", "1e-2")))
res = {}
for path in ("auto", "sparse"):
    eng = Engine()
    eng.profile_path = path
    eng.overlap = False
    eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    eng.profile(tasks, **kw)                       # warm-up
    torch.cuda.synchronize()
    eng.timing = {}
    t0 = time.perf_counter()
    for _ in range(3):
        b = eng.profile(tasks, return_r=(_ == 2), **kw)
    torch.cuda.synchronize()
    ms = eng.collect_timing()
    res[path] = b
    print(f"{path:7s} -> paths {sorted(set(eng.last_paths))} fallbacks {eng.fallbacks}: {1e3 * (time.perf_counter() - t0) / 3:.1f} ms per "
          f"pass of {len(tasks)} tasks; stages (ms per pass): " + ", ".join(f"{k} {v / 3:.2f}" for k, v in ms.items()),
          f"; matches {int(b.hit_counts.sum())} = {b.hit_counts.sum() / (len(tasks) * cfg['N'] * cfg['L']):.3%} of the windows", flush=True)
a, s = res["auto"].r.astype(np.float64), res["sparse"].r.astype(np.float64)
scale = np.abs(s).max(axis=-1, keepdims=True)
print("r tensor vs sparse: max |d| / (1e-5 |r| + 1e-5 rowmax) =", float((np.abs(a - s) / (1e-5 * np.abs(s) + 1e-5 * scale + 1e-12)).max()))

"""Randomised parity check of the CUDA path against the oracle over odd shapes (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth
from mepp_b200.engine import Engine
from oracle import profile, staging

rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "1")))
eng = Engine()
t_end = time.time() + float(os.environ.get("FUZZ_SECONDS", "60"))
n_cases = n_tasks = 0
while time.time() < t_end:
    N = int(rng.integers(4, 300)); L = int(rng.integers(33, 420)); P = int(rng.choice([0, 1, 7, 33, 100, 257]))
    margin = int(rng.choice([0, 0, 1, 2, 2, 5, 16])); use_gc = bool(rng.integers(0, 2))
    path = str(rng.choice(os.environ.get("FUZZ_PATHS", "sparse,tensor,auto").split(',')))
    ascii_u8, scores = synth.synth_sequences(N, L, seed=int(rng.integers(1 << 30)), n_rate=float(rng.choice([0, 0.01, 0.2])),
                                             lower_rate=float(rng.choice([0, 0.02])))
    perm = synth.synth_permutations(N, P, seed=int(rng.integers(1 << 30))) if P else None
    ms = [int(rng.integers(1, min(32, L) + 1)) for _ in range(int(rng.integers(1, 6)))]
    tasks = []
    for m in ms:
        pwm = rng.dirichlet(np.full(4, float(rng.choice([0.05, 0.3, 1.0]))), size=m).T
        for o in rng.choice(['+', '-', '+/-'], size=int(rng.integers(1, 4)), replace=False):
            tasks.append((pwm, str(o)))
    eng.profile_path = path
    eng.stage(ascii_u8, scores, perm, use_gc=use_gc)
    b = eng.profile(tasks, margin=margin, return_hits=True, return_r=True)
    codes = staging.ascii_to_codes(ascii_u8)
    Y = staging.permuted_scores(scores, perm) if P else scores[:, None]
    z = staging.gc_ratio_global(codes) if use_gc else None
    for t, (pwm, o) in enumerate(tasks):
        orc = profile.profile_task(codes, Y, z, pwm, o, margin=margin)
        X0 = b.motif_score_matrix(t, L)
        ok = np.array_equal(X0.view(np.uint32), orc["X0"].view(np.uint32)) and b.thr[t] == orc["thr"]
        d = np.abs(b.r[t].astype(np.float64) - orc["r"].astype(np.float64))
        scale = np.abs(orc["r"]).max(axis=-1, keepdims=True)
        # (the tensor path used to need a floor of 1e-4 here: with one global scale the fp16 hi/lo operands lost relative
        # precision on rows whose scores are all tiny; the per-task scale of mepp_x_scales closed that)
        floor = float(os.environ.get("FUZZ_FLOOR", "1e-5"))
        ok_r = bool((d <= 1e-5 * np.abs(orc["r"]) + floor * scale + 1e-12).all())
        pos_ok = True
        for col in orc["positions"]:
            a_, b_ = np.asarray(b.positions[col][t], dtype=np.float64), np.asarray(orc["positions"][col], dtype=np.float64)
            if col.endswith("_pval") and col.startswith("permutation"):
                pos_ok &= bool(np.abs(a_ - b_).max() <= 0.26)          # discrete: k / P flips on 1-ulp ties
            elif col == "positional_r_pval":
                pos_ok &= bool(np.allclose(a_, b_, rtol=1e-3, atol=1e-9, equal_nan=True))
            else:
                sc = max(np.nanmax(np.abs(b_)) if b_.size else 0.0, 1e-30)
                pos_ok &= bool(np.nanmax(np.abs(a_ - b_)) <= floor * sc + 2e-7) if b_.size else True
        if not (ok and ok_r and pos_ok):
            print(f"MISMATCH N={N} L={L} P={P} margin={margin} gc={use_gc} path={path} task={t} m={pwm.shape[1]} o={o} "
                  f"matches={ok} r={ok_r} positions={pos_ok}", flush=True)
            sys.exit(1)
        n_tasks += 1
    n_cases += 1
print(f"fuzz ok: {n_cases} datasets, {n_tasks} tasks", flush=True)

#!/usr/bin/env python
"""bench.py -- headline benchmark of the MEPP profiling hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

Metric (BASELINE.json): Gbp*motifs/s scanned + profiled = N_seq * L * tasks / seconds, tasks =
motifs x orientations.  A "step" is one pass of the whole hot path (scan -> tcgen05 profile GEMM ->
correlation -> permutation statistics) over every task of the workload on synthetic scored FASTA
of the named shape.  `value` times steps whose inputs are already staged in HBM; `e2e` times the
same step through the public API from pinned HOST buffers: H2D of sequences / scores /
permutation indices, 2-bit pack, score-operand build, host MOODS thresholds, all tasks, and D2H of
every result table.  With --gpus N the ONE task list of the workload is sharded over N ranks, as the
reference fans one list of motif x orientation tasks over its workers (mepp/batch.py:73-140):
strong scaling, the default; `--scaling weak` runs N copies of the list instead.  Every rank stages
its own replica (the upload is split over the ranks and all-gathered over NVLink) and the per-task
tables are gathered to rank 0 with one NCCL collective.

After the timed regions rank 0 checks a few tasks of the workload AT FULL SIZE against the oracle
(`accuracy`: SURVEY.md section 8d's gates) and times the CPU port on a bounded sample (`cpu_baseline`).
`--mode fullrun` times run_mepp itself, wall clock: scored FASTA file -> every task directory.

--impl reference times the CPU baseline (oracle/baseline.py: the reference's algorithm on the
host cores; the reference's own TensorFlow/MOODS code cannot run in this image) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1, which would throttle its BLAS calls
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GATHER_COLUMNS = ['positional_r', 'permutation_positional_r_lower', 'permutation_positional_r_upper',
                  'permutation_positional_r_pval']
METRIC = "Gbp*motifs/s scanned+profiled"
UNIT = "Gbp*motifs/s"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu --set full
# capture of the same bench command (profiles/r02_kernels_ncu_summary.txt); null for workloads without a capture
# (tc_scan4i_kernel: profiles/r02_tc4i_ncu_summary.txt; tc_scan4_kernel, MEPP_TC_ISSUERS=2: profiles/r02_kernels_ncu_summary.txt)
NCU_DRAM_BYTES_PER_LAUNCH = {("tc_scan4i_kernel", "cfg3"): 409.3e6, ("tc_scan4_kernel", "cfg3"): 413.4e6, ("scan_kernel", "cfg2"): 16.0e6}
TC_KERNEL = "tc_scan4_kernel" if os.environ.get("MEPP_TC_ISSUERS", "4") == "2" else "tc_scan4i_kernel"


def clocks_mhz_hint(peaks):
    return float(peaks.get("sm_mhz", 1965.0))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained",
                    p["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def halt(self):
        """Stop polling (NVML polling stalls CUDA API calls: +25 ms per end-to-end step was measured)."""
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2.0)            # really gone before the end-to-end region starts
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.halt()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.2):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


def build_workload(workload, world, scaling="strong"):
    from mepp_b200 import synth
    cfg = synth.make_config(workload)
    base = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    # strong: the ONE task list of the workload, sharded over the ranks (mepp/batch.py:73-140 fans one list over its
    # workers); weak: a copy of the list per GPU
    tasks = base * world if scaling == "weak" else base
    return cfg, tasks


def perm_rows_for_cpu(cfg, max_bytes=4 << 30):
    """Permutation rows for the CPU legs (oracle check, CPU port): the whole matrix, or -- when it only exists as a
    PermutationSource (cfg 5) -- its first rows, as many as fit `max_bytes`.  Returns (rows, all_of_them)."""
    perm = cfg["perm_idx"]
    if not hasattr(perm, "rows"):
        return perm, True
    n = max(64, min(perm.shape[0], (max_bytes // (4 * perm.shape[1])) // 64 * 64))
    return perm.rows(0, n), n >= perm.shape[0]


def bench_config(args, cfg, n_tasks):
    """`config` of the JSON line: the same keys and values for both arms (the driver compares them)."""
    return dict(workload=args.workload, N=cfg["N"], L=cfg["L"], P=cfg["P"], margin=cfg["margin"], tasks_total=n_tasks,
                orientations=",".join(sorted({o for _, _, o in cfg["tasks"]})), scaling=args.scaling,
                l2="no flush needed: every step streams far more than the 126 MB L2 (the correlation matrix r alone is "
                   "4 MB per task, written and re-read by the statistics)")


def run_reference(args, rank, world):
    """CPU baseline arm: rank 0 only; other ranks exit 0 without work."""
    if rank != 0:
        return
    from oracle import baseline, staging
    cfg, tasks = build_workload(args.workload, world, args.scaling)
    codes = staging.ascii_to_codes(cfg["ascii"])
    perm_cpu, perm_all = perm_rows_for_cpu(cfg)
    Y = staging.permuted_scores(cfg["scores"], perm_cpu)
    z = staging.gc_ratio_global(codes)
    cores = os.cpu_count() or 1
    # a step = a bounded sample of the workload: whole tasks until the step's share of ~2 minutes is spent (at least one)
    per_step_budget = min(20.0, 120.0 / max(1, args.steps + args.warmup))
    n_done = 0
    for _ in range(args.warmup):
        baseline.time_sample(codes, Y, z, tasks, cfg["margin"], threads=cores, budget_s=per_step_budget / 2)
    total_t, total_tasks = 0.0, 0
    for _ in range(args.steps):
        t, n_done = baseline.time_sample(codes, Y, z, tasks, cfg["margin"], threads=cores, budget_s=per_step_budget)
        total_t += t
        total_tasks += n_done
    value = cfg["N"] * cfg["L"] * total_tasks / total_t / 1e9
    sample = (f"{total_tasks} whole tasks of {args.workload} (first {n_done} motif x orientation tasks per step, "
              f"N={cfg['N']} L={cfg['L']} P={Y.shape[1] - 1}{'' if perm_all else ' of ' + str(cfg['P'])}), "
              f"float32 scan in C + BLAS sgemm profile + numpy stats")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total_t / max(1, args.steps), higher_is_better=True,
                scaling=args.scaling, vs_baseline=None, dtype="f32", data="synthetic",
                config=bench_config(args, cfg, len(tasks)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def accuracy_block(eng, cfg, my_tasks, my_ids, n_check, margin):
    """SURVEY.md 8(d) accuracy gates on a few of this rank's tasks at FULL size: the device results of the default
    path (the kernels the timed steps ran) against the oracle."""
    from oracle import checks, staging
    n_loc = len(my_tasks)
    if n_loc == 0 or n_check <= 0:
        return None
    # the first motif in both orientations (planted signal), one from the middle, the last one
    pick = sorted({0, min(1, n_loc - 1), (n_loc // 2) & ~1, n_loc - 1})[:max(1, n_check)]
    sel = [my_tasks[i] for i in pick]
    t0 = time.perf_counter()
    got = eng.profile(sel, margin=margin, return_hits=True, return_r=True)
    hits = got.hits.copy()
    codes = staging.ascii_to_codes(cfg["ascii"])
    perm_cpu, perm_all = perm_rows_for_cpu(cfg)
    Y = staging.permuted_scores(cfg["scores"], perm_cpu)
    z = staging.gc_ratio_global(codes)
    # (when only the first permutation rows are checked, the statistics over ALL permutations are not compared)
    rep = checks.accuracy_report(codes, Y, z, sel, [my_ids[i] for i in pick], margin, hits, got.r[:, :, :Y.shape[1]],
                                 got.positions if perm_all else None)
    rep["permutation_columns_checked"] = int(Y.shape[1] - 1)
    rep["scan_form"] = getattr(eng, "last_scan", None)
    rep["profile_path"] = ",".join(sorted(set(eng.last_paths)))
    rep["seconds"] = round(time.perf_counter() - t0, 1)
    return rep


def write_fullrun_inputs(cfg, workdir):
    """Scored FASTA (`>id score` headers) + motif file of the workload, as the `mepp` CLI reads them."""
    fa = os.path.join(workdir, "scored.fa")
    N, L = cfg["ascii"].shape
    rows = np.ascontiguousarray(cfg["ascii"]).view(f"S{L}").ravel()
    with open(fa, "w") as f:
        for n in range(N):
            f.write(f">seq{n:07d} {float(cfg['scores'][n]):.6f}\n")
            f.write(rows[n].decode("latin-1"))
            f.write("\n")
    mf = os.path.join(workdir, "motifs.txt")
    seen = set()
    with open(mf, "w") as f:
        for key, pwm, _o in cfg["tasks"]:
            if key in seen:
                continue
            seen.add(key)
            f.write(f">{key}\n")
            for c in range(4):
                f.write("\t".join(repr(float(v)) for v in pwm[c]) + "\n")
    return fa, mf


def run_fullrun(args, rank, world, cfg):
    """The metric's second half: wall clock of run_mepp itself -- scored FASTA file and motif file in, the complete
    output directory out (filtered_data_df, one directory per motif x orientation with positions_df / ranks_df /
    yaml files, filepaths.tsv, results tables) -- at `world` GPUs."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from mepp_b200 import mepp as mepp_mod
    workdir = None
    if rank == 0:
        workdir = tempfile.mkdtemp(prefix="mepp_fullrun_")
        write_fullrun_inputs(cfg, workdir)
    if world > 1:
        box = [workdir]
        dist.broadcast_object_list(box, src=0)
        workdir = box[0]
    fa, mf, out = os.path.join(workdir, "scored.fa"), os.path.join(workdir, "motifs.txt"), os.path.join(workdir, "out")
    orients = sorted({o for _, _, o in cfg["tasks"]}, key=["+", "-", "+/-"].index)
    runs = []
    for it in range(1 + max(0, args.fullrun_repeats)):          # the first run is the warm-up (CUDA context, page cache)
        if rank == 0:
            shutil.rmtree(out, ignore_errors=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fdf, _, _ = mepp_mod.run_mepp(fa, mf, out, num_permutations=cfg["P"], motif_margin=cfg["margin"],
                                      motif_orientations=orients, permutation_seed=cfg_seed(cfg) + 1, shuffle_seed=1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        runs.append(dict(wall_s=wall, phases={k: round(v, 3) for k, v in mepp_mod.LAST_RUN_TIMINGS.items()},
                         failed_tasks=int((fdf["retval"] != 0).sum()), tasks=int(len(fdf))))
    n_files = n_bytes = 0
    if rank == 0:
        for d, _dirs, files in os.walk(out):
            for fn in files:
                n_files += 1
                n_bytes += os.path.getsize(os.path.join(d, fn))
        shutil.rmtree(workdir, ignore_errors=True)
    best = min(runs[1:] or runs, key=lambda r: r["wall_s"])
    return dict(wall_s=best["wall_s"], runs=[round(r["wall_s"], 3) for r in runs], phases=best["phases"],
                tasks=best["tasks"], failed_tasks=best["failed_tasks"], files_written=n_files, bytes_written=n_bytes,
                gbp_motifs_per_s=cfg["N"] * cfg["L"] * best["tasks"] / best["wall_s"] / 1e9,
                what="run_mepp(scored FASTA file, motif file, out dir): parse, filter, order, permutations, staging, "
                     "every task, every task directory, filepaths.tsv, results tables; wall clock, max over ranks; "
                     "the first run is a warm-up")


def cfg_seed(cfg):
    from mepp_b200 import synth
    return synth.SEED_BASE + int(cfg["cfg"][3:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", help="cfg1 .. cfg5 of BASELINE.json (default: cfg3, the north-star target)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the workload's one task list sharded over the GPUs; weak: one copy of the list per GPU")
    ap.add_argument("--mode", default="step", choices=["step", "fullrun"],
                    help="step: the hot path per step (the JSON line of the contract); fullrun: run_mepp wall clock, "
                         "scored FASTA file -> complete output directory")
    ap.add_argument("--fullrun-repeats", type=int, default=1, help="timed run_mepp runs after the warm-up run")
    ap.add_argument("--no-fullrun", action="store_true", help="skip the run_mepp wall-clock measurement of the step line")
    ap.add_argument("--simulate-world", type=int, default=0,
                    help="diagnostic: one process runs only shard 0 of this many shards of the task list (the per-rank "
                         "work of a strong-scaling run); the printed value is NOT a bench value")
    ap.add_argument("--no-accuracy", action="store_true", help="skip the full-size oracle check of a few tasks")
    ap.add_argument("--accuracy-tasks", type=int, default=4)
    ap.add_argument("--no-tensor-leg", action="store_true",
                    help="skip the extra tensor-path (tcgen05 profile GEMM) measurement on a slice of the task list")
    ap.add_argument("--tensor-leg-tasks", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not run the nvidia-smi clock sampler")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from mepp_b200 import parallel, motif_layers, build
    build.build()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # (NCCL's version banner: stdout carries ONE JSON line)
    rank, world, local = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mepp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from mepp_b200.engine import Engine
    import torch.distributed as dist

    cfg, tasks = build_workload(args.workload, world, args.scaling)
    N, L, P, margin = cfg["N"], cfg["L"], cfg["P"], cfg["margin"]
    if args.mode == "fullrun":
        fr = run_fullrun(args, rank, world, cfg)
        if rank == 0:
            n_tasks = fr["tasks"]
            print(json.dumps(dict(metric="full MEPP run wall-time", value=fr["wall_s"], unit="s", n_gpus=world,
                                  steps=max(1, args.fullrun_repeats), warmup=1, ms_per_step=1e3 * fr["wall_s"],
                                  higher_is_better=False, scaling="strong", vs_baseline=None, dtype="f32",
                                  data="synthetic", config=bench_config(args, cfg, n_tasks), fullrun=fr)), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    T_total = len(tasks)
    my_ids = parallel.shard_task_units(tasks, world)[rank]    # LPT over (motif) units: orientation pairs stay together
    if args.simulate_world > 1 and world == 1:
        my_ids = parallel.shard_task_units(tasks, args.simulate_world)[0]
    my_tasks = [tasks[i] for i in my_ids]

    # pinned host inputs (the e2e path copies them every step)
    h_ascii = torch.from_numpy(cfg["ascii"]).pin_memory()
    h_scores = torch.from_numpy(cfg["scores"]).pin_memory()
    lazy_perm = hasattr(cfg["perm_idx"], "rows")      # cfg 5: the 40 GB index matrix is generated and staged in chunks
    h_perm = cfg["perm_idx"] if lazy_perm else torch.from_numpy(cfg["perm_idx"]).pin_memory()
    # host -> device bytes of one end-to-end step on this rank (sharded upload + all-gather when world > 1)
    h2d_bytes = (-(-h_ascii.shape[0] // world) * h_ascii.shape[1] + 4 * h_scores.numel()
                 + 4 * -(-h_perm.shape[0] // world) * h_perm.shape[1] + len(my_tasks) * (32 * 4 * 2 * 4 + 12))

    eng = Engine(device=local)

    phase_ms = []

    def e2e_step():
        """Public-API step from host buffers: stage + thresholds + all tasks (+ gather to rank 0 when sharded)."""
        dbg = bool(os.environ.get("MEPP_BENCH_DEBUG"))
        tq = [time.perf_counter()]
        if world > 1:
            # every rank uploads 1 / world of the sequences and permutation indices, all-gathered over NVLink
            eng.stage_allgather(h_ascii, h_scores, h_perm, use_gc=True)
        else:
            eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
        tq.append(time.perf_counter())
        batch = eng.profile(my_tasks, margin=margin)       # thresholds on the device, groups pipelined
        tq.append(time.perf_counter())
        full = None
        if world > 1:
            # the small per-task profiles go to rank 0 (float32, one row per column; the full tables stay per rank,
            # as run_batch writes each rank's task directories locally)
            block = np.empty((len(my_tasks), len(GATHER_COLUMNS), L), dtype=np.float32)
            for j, k in enumerate(GATHER_COLUMNS):
                block[:, j, :] = batch.positions[k]
            full = parallel.gather_blocks(block, my_ids, T_total, device=dev)
        if dbg:
            tq.append(time.perf_counter())
            phase_ms.append([round(1e3 * (b - a), 2) for a, b in zip(tq[:-1], tq[1:])])
        return batch, full

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- warm-up (e2e style so that every buffer exists), then the two timed regions
    for _ in range(max(args.warmup, 1)):
        batch, full = e2e_step()
    tables = motif_layers.scan_tables(my_tasks)
    # bytes actually copied device -> host per step: the finalised column block (68 B per (task, position), 44 B
    # per task), density per (task, sequence), the match / entry counters of every group
    n_loc = len(my_tasks)
    d2h_bytes = int(n_loc * L * 68 + n_loc * 44 + n_loc * eng.n_pad * (2 if L <= 65535 else 4)
                    + eng.last_groups * (8 + 8 * 16 * 64))

    sampler = ClockSampler(local)
    if not args.no_clocks:
        sampler.start()
    time.sleep(0.3)

    # (1) device-resident: inputs staged in HBM before the timed region
    eng.stage(h_ascii, h_scores, h_perm, use_gc=True)
    dev_tables = eng.upload_tables(tables)               # device-resident: sequences, scores, permutations and tables
    eng.profile(my_tasks, margin=margin, tables=dev_tables)
    sync_all()
    import ctypes as C
    from mepp_b200 import _lib
    _lib.check(_lib.lib.mepp_gemm_counters(None, 1), "gemm_counters")
    launches0 = eng.launches
    # timed on the device: CUDA events on the launching stream, a barrier + synchronize on both sides (every
    # profile() call ends with a host wait for its last result copy, so the closing event is recorded after all of
    # the step's work on the side streams); the host clock around the same region is kept as a cross-check
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    # a stream of steps through the asynchronous form of the call: the host side of step k's last task group (wait for
    # its result copies, move them into step k's tables) runs while step k + 1 is on the GPU; every step's complete
    # result tables are on the host when its result() returns, and the last one inside the timed region
    eng.host_timing = {}
    ev_a.record(torch.cuda.current_stream(dev))
    prev = None
    for _ in range(args.steps):
        cur = eng.profile_async(my_tasks, margin=margin, tables=dev_tables)
        if prev is not None:
            prev.result()
        prev = cur
    prev.result()
    ev_b.record(torch.cuda.current_stream(dev))
    host_phases = {k: round(v / max(1, args.steps), 3) for k, v in eng.host_timing.items() if k != 'calls'}
    eng.host_timing = None
    sync_all()
    t1 = time.perf_counter()
    wall_value = t1 - t0
    launches = eng.launches - launches0
    # per-kernel CUDA-event times: a serial pass (no overlap of the scan with the previous group's chain, so that
    # every stage is timed alone on the launching stream) over the same inputs, right after the timed region
    kernel_steps = max(1, min(args.steps, 5))
    overlap_was = eng.overlap
    eng.overlap = False
    eng.profile(my_tasks, margin=margin, tables=dev_tables)
    sync_all()
    eng.timing = {}
    _lib.check(_lib.lib.mepp_scan_tc_timing(1, None, None, None, None), "scan_tc_timing")   # events around tc_scan_kernel
    for _ in range(kernel_steps):
        eng.profile(my_tasks, margin=margin, tables=dev_tables)
    sync_all()
    stage_ms = eng.collect_timing()
    tc_ms, tc_macs, tc_bytes, tc_n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
    _lib.check(_lib.lib.mepp_scan_tc_timing(0, C.byref(tc_ms), C.byref(tc_macs), C.byref(tc_bytes), C.byref(tc_n)),
               "scan_tc_timing")
    eng.timing = None
    eng.overlap = overlap_was
    steps2 = (C.c_ulonglong * 2)()
    _lib.check(_lib.lib.mepp_gemm_counters(steps2, 0), "gemm_counters")
    issued_frac = (steps2[0] / steps2[1]) if steps2[1] else 1.0
    # the tcgen05 profile GEMM (the path of a dense X: Engine.choose_path) on a slice of the same task list, so that its
    # tensor-pipe figures are in every bench line although the default step takes the sparse path at p = 1e-4
    tensor_leg = None
    if not args.no_tensor_leg and all(p_ == "sparse" for p_ in eng.last_stats["paths"]):
        sub = my_tasks[:min(len(my_tasks), args.tensor_leg_tasks)]
        path_was, eng.profile_path, eng.overlap = eng.profile_path, "tensor", False
        try:
            eng.profile(sub, margin=margin)                      # warm-up: operand buffers
            sync_all()
            _lib.check(_lib.lib.mepp_gemm_counters(None, 1), "gemm_counters")
            eng.timing = {}
            for _ in range(2):
                eng.profile(sub, margin=margin)
            sync_all()
            tl_ms = eng.collect_timing()
            eng.timing = None
            _lib.check(_lib.lib.mepp_gemm_counters(steps2, 0), "gemm_counters")
            tl_frac = (steps2[0] / steps2[1]) if steps2[1] else 1.0
            t_gemm = tl_ms.get("gemm", 0.0) * 1e-3 / 2
            if t_gemm > 0:
                alg = 2.0 * L * N * (1 + P) * len(sub)
                ex = 3.0 * alg * tl_frac * (eng.ldS / float(1 + P))
                tensor_leg = dict(tasks=len(sub), bound="tensor", achieved=ex / t_gemm / 1e12, unit="TFLOP/s",
                                  peak=None, frac=None, algorithmic_tflops=alg / t_gemm / 1e12, issued_step_fraction=tl_frac,
                                  split_products=3, ms=1e3 * t_gemm, paths=sorted(set(eng.last_stats["paths"])),
                                  note="profile_gemm_kernel (tcgen05 kind::f16, fp16 hi/lo split, 3 products, fp32 TMEM "
                                       "accumulators, correlation epilogue fused) on the first tasks of the list; executed = "
                                       "issued MMAs only (all-zero K = 16 steps of X are skipped)")
        finally:
            eng.profile_path, eng.overlap, eng.timing = path_was, overlap_was, None
    dt_value = parallel.barrier_and_max(1e-3 * ev_a.elapsed_time(ev_b), device=dev if world > 1 else None)

    # (2) end to end from pinned host buffers (clock sampler halted: its NVML polling slows host-side CUDA calls)
    sampler.halt()
    import gc
    gc.collect()
    gc.disable()                                           # no collector pauses inside the timed steps
    for _ in range(max(args.warmup, 3)):                   # untimed warm-up of the end-to-end path (the first steps after
        e2e_step()                                         # the switch from the device-resident loop are erratic)
    sync_all()
    ev_c, ev_d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2 = time.perf_counter()
    ev_c.record(torch.cuda.current_stream(dev))
    e2e_times = []
    for _ in range(args.steps):
        ta = time.perf_counter()
        e2e_step()
        e2e_times.append(round(1e3 * (time.perf_counter() - ta), 1))
    ev_d.record(torch.cuda.current_stream(dev))
    sync_all()
    t3 = time.perf_counter()
    gc.enable()
    if os.environ.get("MEPP_BENCH_DEBUG"):
        print(f"rank {rank} e2e host ms per step:", e2e_times, "phases (stage, profile, gather):",
              phase_ms[-args.steps:], file=sys.stderr, flush=True)
    # an end-to-end step ends on the host (results in the caller's arrays, gathered on rank 0), so the larger of the
    # event time and the host clock is what a caller waits for
    dt_e2e = parallel.barrier_and_max(max(1e-3 * ev_c.elapsed_time(ev_d), t3 - t2), device=dev if world > 1 else None)
    clocks = sampler.stop(t0, t1)
    clocks["sampled"] = "during the device-resident timed region (100 ms period); halted for the e2e region"

    # (3) the metric's second half: run_mepp wall clock, files in -> output directory out (every rank takes part)
    fullrun = None
    if not args.no_fullrun and args.simulate_world <= 1:
        fullrun = run_fullrun(args, rank, world, cfg)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    work_step = float(N) * L * T_total                      # bp x tasks per step, all ranks
    value = work_step * args.steps / dt_value / 1e9
    e2e_value = work_step * args.steps / dt_e2e / 1e9
    n_local = len(my_tasks)
    groups = eng.last_groups
    # per-kernel rooflines from the CUDA-event stage times of the timed value region (rank 0's share).
    # Algorithmic bytes / flops per launch group are stated in DESIGN.md section 4; the counts come from the
    # engine (matches, non-zeros of X, scan units of the last step).
    C_ = 1 + P
    st = eng.last_stats
    sparse = all(p_ == "sparse" for p_ in st["paths"])
    steps = kernel_steps
    per = lambda k: stage_ms[k] * 1e-3 / steps             # seconds per step (serial pass)
    kernels = {}
    l2_cap_gbs = 6300.0 * clocks_mhz_hint(peaks) / 1e3     # LTS throughput cap, B300_MICROARCH.md: ~6300 B/clk full chip
    if stage_ms.get("scan"):
        # reads 3 bits per base per scan unit (a '+/-' task and its '+' twin share one pass), writes 16 B per match and
        # 16 B per non-zero of X (sparse path) or 32 B per non-zero 8-sequence piece (tiled path, <= one per non-zero)
        lean = bool(getattr(eng, "lean", False)) and sparse       # matches also go to buckets (8 B each), read back once
        tc = getattr(eng, "last_scan", None) == "tc"
        scan_bytes = (0.375 * N * L * st["units"] + (32.0 if lean else 16.0) * st["hits"]
                      + (16.0 if sparse else 32.0) * max(st["entries"], st["hits"]))
        if tc:        # the first stage reads the flat one-hot stream (4 B per base) once per launch instead
            scan_bytes += tc_bytes.value / steps - 0.375 * N * L * st["units"]
        a = scan_bytes / per("scan") / 1e9
        kernels["scan"] = dict(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"],
                               ms_per_step=1e3 * per("scan"),
                               launches_per_step=(tc_n.value // steps) * 3 + groups if tc else groups * (2 if lean else 1),
                               form="tc: tcgen05 int8 first stage over the one-hot stream (tc_scan_kernel), exact float32 "
                                    "evaluation of its candidates (tc_exact_kernel), row sums from per-(task, k-block) "
                                    "buckets of matches (bucket_entries_kernel)" if tc else
                                    "lean: matches only + entries built from per-(task, k-block) buckets" if lean
                                    else "full: blur and entries from staging rows inside the scan",
                               gwindows_per_s=N * L * n_local / per("scan") / 1e9,
                               # a scan that writes X once moves 4 B per bp*task; this one emits X as a list of
                               # its non-zeros, so per 8(d) it is reported as compute bound: adds = N * L * m * strands
                               survey_4B_equiv_gbs=4.0 * N * L * n_local / per("scan") / 1e9,
                               survey_4B_equiv_frac=4.0 * N * L * n_local / per("scan") / 1e9 / peaks["hbm_gbs"],
                               tap_adds_per_s=float(N) * L * sum(pwm.shape[1] * (1 if o in ('+', '-') else 2)
                                                                 for pwm, o in my_tasks) / per("scan"),
                               int_add_peak_per_s=148 * 128 * clocks_mhz_hint(peaks) * 1e6,
                               limit=("tensor pipe hand-overs of the first stage (tc_scan_kernel, see roofline) + atomics of "
                                      "the match publication; X is written as a list of its matches, so almost no bytes move")
                                     if tc else
                                     "integer issue / shared-memory lookups of the pre-filter (ncu: issue slots ~47 % busy, "
                                     "DRAM < 1 %): X is written as a list of its non-zeros, so almost no bytes move")
        if tc and tc_ms.value > 0:
            t_tc = tc_ms.value * 1e-3 / steps                          # seconds of tc_scan_kernel per step (CUDA events)
            tops = 2.0 * tc_macs.value / steps / t_tc / 1e12
            i8_peak = 2.0 * peaks["tf_sustained"]
            alg = 2.0 * float(N) * L * sum(pwm.shape[1] * (1 if o in ('+', '-') else 2) for pwm, o in my_tasks)
            kernels["tc_scan_kernel"] = dict(
                bound="tensor", achieved=tops, peak=i8_peak, unit="TOP/s", frac=tops / i8_peak,
                ms_per_step=1e3 * t_tc, launches_per_step=tc_n.value // steps, ms_per_launch=tc_ms.value / max(1, tc_n.value),
                peak_note="dense int8 = 2 x the measured sustained bf16 rate of MEASURED_PEAKS.json (kind::i8 issues K = 32 "
                          "per instruction against 16 for bf16); burst: %.0f" % (2.0 * peaks["tf_burst"]),
                issued_note="issued multiply-adds: 128 windows x N-tile columns x 32 channels per K-step, the padding of "
                            "the K-steps (taps + 3 phases rounded up to 8 bases) and of the N-tiles included",
                algorithmic_tops=alg / t_tc / 1e12,
                algorithmic_note="SURVEY 8(d) compute-bound form: N * L * m * strands multiply-adds per task",
                stream_gbs=tc_bytes.value / steps / t_tc / 1e9,
                kernel=TC_KERNEL,
                ncu=("profiles/r02_tc4i_ncu_summary.txt (cfg3, tc_scan4i_kernel): tensor pipe 44 % active, TMEM port 58 % busy, issue "
                     "slots 56 % busy, DRAM 409 MB per launch (the one-hot stream, read once)" if TC_KERNEL == "tc_scan4i_kernel" else
                     "profiles/r02_kernels_ncu_summary.txt (cfg3, tc_scan4_kernel): tensor pipe 20 % active, issue slots 53 % "
                     "busy, DRAM 413 MB per launch (the one-hot stream, read once)"),
                limit="the length of the hand-over chain of an accumulator stage (issuer wake-up, tcgen05.mma issue, "
                      "completion, epilogue wake-up, TMEM reads, release: four stages of 128 TMEM columns, one issuer and "
                      "one epilogue group per stage) and the instruction stream of the epilogue warps; floors: TMEM reads "
                      "at 256 B/clk/SM ~2.1 ms per cfg3 step, products ~1.2 ms: DESIGN.md section 4")
    if stage_ms.get("gemm") and sparse:
        # HBM: Y rows once (L2-resident afterwards), CSR records, r written once; L2 -> SM: one Y row per non-zero
        ldY = -(-C_ // 512) * 512
        by_match = bool(getattr(eng, "sparse_matches", False))
        y_bytes = 4.0 * eng.n_pad * ldY
        gather = 4.0 * ldY * st["entries"]               # one float32 row of Y per record (match, or non-zero of X)
        if by_match:
            # Y rows gathered once per match (from HBM when Y exceeds the L2: cfg 3), S0 float64 written and read once,
            # the match CSR, r written once
            hbm = (gather if y_bytes > 100e6 else y_bytes) + 2 * 8.0 * n_local * L * ldY + 24.0 * st["entries"] \
                  + 4.0 * n_local * L * eng.ldS
        else:
            hbm = y_bytes + 24.0 * st["entries"] + 4.0 * n_local * L * eng.ldS
        a = hbm / per("gemm") / 1e9
        kernels["profile_sparse"] = dict(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s",
                                         frac=a / peaks["hbm_gbs"], ms_per_step=1e3 * per("gemm"),
                                         launches_per_step=(4 if by_match else 3) * groups,
                                         form="matches: S0 = one gathered row of Y per match (float64), window sums of S0 + "
                                              "correlation epilogue" if by_match else
                                              "entries: one gathered row of Y per non-zero of the blurred matrix",
                                         gather_gbs=gather / per("gemm") / 1e9,
                                         l2_cap_gbs=l2_cap_gbs, gather_frac_of_l2_cap=gather / per("gemm") / 1e9 / l2_cap_gbs,
                                         records_per_step=st["entries"],
                                         limit="gather of one float32 Y row (4 * ldY bytes) per record: HBM when Y (%.0f MB) "
                                               "exceeds the L2, else L2 -> SM" % (y_bytes / 1e6))
    elif stage_ms.get("gemm"):
        # executed = issued MMAs only (all-zero K=16 steps of X are skipped); padded tile columns are counted as issued
        gemm_alg = 2.0 * L * N * C_ * n_local
        padded = (eng.ldS / float(C_))
        a = 3.0 * gemm_alg * issued_frac * padded / per("gemm") / 1e12
        kernels["profile_gemm"] = dict(bound="tensor", achieved=a, peak=peaks["tf_sustained"], unit="TFLOP/s",
                                       frac=a / peaks["tf_sustained"], frac_of_burst=a / peaks["tf_burst"],
                                       algorithmic_tflops=gemm_alg / per("gemm") / 1e12,
                                       issued_step_fraction=issued_frac, split_products=3,
                                       ms_per_step=1e3 * per("gemm"), launches_per_step=groups)
    if stage_ms.get("perm_stats"):
        # r is streamed three times: row selection, pooled histogram + second radix digit, integrals + third digit
        stats_bytes = 4.0 * n_local * L * 3.0 * C_
        a = stats_bytes / per("perm_stats") / 1e9
        kernels["perm_stats"] = dict(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"],
                                     ms_per_step=1e3 * per("perm_stats"), launches_per_step=9 * groups,
                                     limit="per-row tail selection in registers, pooled histogram search, 3-digit radix select: ALU / shared-memory atomics")
    for k in ("correlation", "finalize"):
        if stage_ms.get(k):
            kernels[k] = dict(ms_per_step=1e3 * per(k))
    if tensor_leg is not None:
        tensor_leg["peak"] = peaks["tf_sustained"]
        tensor_leg["frac"] = tensor_leg["achieved"] / peaks["tf_sustained"]
        tensor_leg["frac_of_burst"] = tensor_leg["achieved"] / peaks["tf_burst"]
        kernels["profile_gemm"] = tensor_leg
    # the dominant kernel of the step (the longest single kernel of the ncu launch list, profiles/)
    if "tc_scan_kernel" in kernels:
        dk = kernels["tc_scan_kernel"]
        roofline = dict(kernel=TC_KERNEL, bound="tensor", achieved=dk["achieved"], peak=dk["peak"], unit=dk["unit"],
                        frac=dk["frac"], traffic=NCU_DRAM_BYTES_PER_LAUNCH.get((TC_KERNEL, args.workload)),
                        peak_source=peaks["source"] + "; " + dk["peak_note"], ms_per_launch=dk["ms_per_launch"],
                        algorithmic=dk["algorithmic_tops"], share_of_step=dk["ms_per_step"] / (1e3 * dt_value / args.steps),
                        note="achieved = int8 multiply-adds issued by the launches x 2 / their CUDA-event time (events on the "
                             "launching stream around every " + TC_KERNEL + " launch of a serial pass, mepp_scan_tc_timing); "
                             "algorithmic = 2 N L m strands per task / the same time; traffic = dram__bytes_read + write of "
                             "one launch, ncu --set full of this command (profiles/r02_tc4i_ncu_summary.txt)")
    else:
        dominant = "scan" if "scan" in kernels else next(iter(kernels))
        dk = kernels.get(dominant, {})
        roofline = dict(kernel=dominant + "_kernel", bound=dk.get("bound"), achieved=dk.get("achieved"), peak=dk.get("peak"),
                        unit=dk.get("unit"), frac=dk.get("frac"),
                        traffic=NCU_DRAM_BYTES_PER_LAUNCH.get((dominant + "_kernel", args.workload)),
                        peak_source=peaks["source"],
                        note="achieved = algorithmic bytes / CUDA-event time of the launches (DESIGN.md section 4)")

    accuracy = None
    if not args.no_accuracy:
        accuracy = accuracy_block(eng, cfg, my_tasks, my_ids, args.accuracy_tasks, margin)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import baseline, staging
        codes = staging.ascii_to_codes(cfg["ascii"])
        Y = staging.permuted_scores(cfg["scores"], perm_rows_for_cpu(cfg)[0])
        z = staging.gc_ratio_global(codes)
        cores = os.cpu_count() or 1
        t, n_done = baseline.time_sample(codes, Y, z, tasks, margin, threads=cores, budget_s=15.0)
        cpu_baseline = dict(value=N * L * n_done / t / 1e9, unit=UNIT, cores=cores, kind="port",
                            sample=f"first {n_done} tasks of {args.workload} (N={N}, L={L}, P={Y.shape[1] - 1}) in {t:.1f} s; "
                                   f"float32 scan in C + BLAS sgemm profile + numpy statistics")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 1),
                ms_per_step=1e3 * dt_value / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f32", data="synthetic",
                config=bench_config(args, cfg, T_total),
                details=dict(tasks_per_gpu=n_local, parallelism=f"tasks sharded over {world} GPU(s)",
                             scan_form=getattr(eng, "last_scan", None),
                             profile_path=",".join(sorted(set(st["paths"]))),
                             profile="sparse: float64 gather of float32 Y rows per non-zero of X; "
                                     "tensor: fp16 hi/lo split, 3 tcgen05 products, fp32 TMEM accumulate",
                             fallbacks=list(eng.fallbacks), host_ms_per_step=host_phases),
                e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=1e3 * dt_e2e / args.steps,
                         h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes)),
                timing=dict(how="CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks",
                            host_clock_ms_per_step=1e3 * wall_value / args.steps,
                            overlap="scan of group g+1 on the compute stream, profile / statistics / finalisation of "
                                    "group g on a high-priority stream" if eng.overlap else "none",
                            kernels="per-kernel times: CUDA events on the launching stream in a serial pass of %d steps "
                                    "(overlap off) right after the timed region" % kernel_steps),
                gpu_launches=int(launches), roofline=roofline, kernels=kernels, accuracy=accuracy, fullrun=fullrun,
                cpu_baseline=cpu_baseline, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

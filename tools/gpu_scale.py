"""Scale check on the GPU: per-stage times and accuracy of the tcgen05 GEMM at full K (run under gpurun)."""
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from mepp_b200 import _lib, synth, motif_layers  # noqa: E402
from mepp_b200.engine import Engine, _ptr  # noqa: E402
from oracle import staging, profile  # noqa: E402


def main():
    cfgname = os.environ.get("SCALE_CFG", "cfg2")
    n_motifs = int(os.environ.get("SCALE_MOTIFS", "0")) or None
    n_oracle = int(os.environ.get("SCALE_ORACLE", "2"))
    t0 = time.time()
    cfg = synth.make_config(cfgname, n_motifs=n_motifs, n_seq=int(os.environ.get("SCALE_N", "0")) or None,
                            n_perm=int(os.environ.get("SCALE_P", "0")) or None)
    print(f"synth {cfgname}: N={cfg['N']} L={cfg['L']} P={cfg['P']} tasks={len(cfg['tasks'])} in {time.time()-t0:.1f}s",
          flush=True)
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    eng = Engine()
    t0 = time.time()
    eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    torch.cuda.synchronize()
    print(f"stage: {time.time()-t0:.3f}s  group_size={eng.group_size()}", flush=True)
    t0 = time.time()
    tables = motif_layers.scan_tables(tasks)
    print(f"tables: {time.time()-t0:.3f}s", flush=True)
    for rep in range(3):
        eng.timing = {}
        t0 = time.time()
        batch = eng.profile(tasks, margin=cfg["margin"], tables=tables)
        torch.cuda.synchronize()
        wall = time.time() - t0
        tm = eng.collect_timing()
        work = cfg["N"] * cfg["L"] * len(tasks)
        flops = 2.0 * cfg["L"] * cfg["N"] * (1 + cfg["P"]) * len(tasks)
        print(f"rep {rep}: wall {wall:.3f}s  stages(ms) { {k: round(v, 2) for k, v in tm.items()} }  "
              f"{work / wall / 1e9:.2f} Gbp*motifs/s; gemm alg {flops / (tm['gemm'] * 1e-3) / 1e12:.1f} TF/s, "
              f"x3 = {3 * flops / (tm['gemm'] * 1e-3) / 1e12:.1f} TF/s; scan {work * 4 / (tm['scan'] * 1e-3) / 1e9:.0f} GB/s",
              flush=True)
    # accuracy: a few tasks against the float64 oracle, and the last group's S against the check kernel
    codes = staging.ascii_to_codes(cfg["ascii"])
    Y = staging.permuted_scores(cfg["scores"], cfg["perm_idx"])
    z = staging.gc_ratio_global(codes)
    sel = list(range(min(n_oracle, len(tasks))))
    if not sel:
        print("DONE (no oracle comparison)", flush=True)
        return
    b2 = eng.profile([tasks[i] for i in sel], margin=cfg["margin"], return_r=True, return_hits=True)
    for j, i in enumerate(sel):
        t0 = time.time()
        orc = profile.profile_task(codes, Y, z, tasks[i][0], tasks[i][1], margin=cfg["margin"])
        d = np.abs(b2.r[j].astype(np.float64) - orc['r'].astype(np.float64))
        tol = 1e-5 * np.abs(orc['r']) + 1e-7
        X0g = b2.motif_score_matrix(j, cfg["L"])
        print(f"task {i}: hits {int((orc['X0']>0).sum())} bit-exact {np.array_equal(X0g.view(np.uint32), orc['X0'].view(np.uint32))}; "
              f"r max abs err {d.max():.3e} (|r| max {np.abs(orc['r']).max():.3e}); >1e-5rel+1e-7abs: {(d>tol).sum()} of {d.size}; "
              f"max rel err where |r|>1e-3: {(d/np.maximum(np.abs(orc['r']),1e-30))[np.abs(orc['r'])>1e-3].max():.3e} "
              f"(oracle {time.time()-t0:.1f}s)", flush=True)
        for col in ('permutation_full_positional_r_lower', 'permutation_semi_positional_r_pval',
                    'permutation_full_positional_r_pval', 'permutation_integral_r_pval', 'integral_r',
                    'permutation_semi_positional_r_upper', 'positional_r_lower', 'smoothed_count'):
            a, b = np.asarray(b2.positions[col][j], dtype=np.float64), np.asarray(orc['positions'][col], dtype=np.float64)
            print(f"     {col}: max abs diff {np.abs(a-b).max():.3e}", flush=True)
    # S vs check kernel on selected rows: rerun the same tasks with the unfused GEMM so that S is materialised
    eng.fuse_correlation = False
    eng.profile([tasks[i] for i in sel], margin=cfg["margin"])
    eng.fuse_correlation = True
    m_rows = len(sel) * cfg["L"]
    rows = np.unique(np.random.default_rng(0).integers(0, m_rows, 64)).astype(np.int32)
    rows_d = torch.from_numpy(rows).cuda()
    Sd = torch.empty((len(rows), eng.C), dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib.mepp_profile_gemm_check(_ptr(eng._bufs['x_planes']), _ptr(eng.y_planes), _ptr(Sd), m_rows,
                                                eng.C, eng.n_pad, _ptr(rows_d), len(rows), None))
    torch.cuda.synchronize()
    S = eng._bufs['S'][:m_rows * eng.ldS].view(m_rows, eng.ldS)[rows_d.long(), :eng.C].double()
    err = (S - Sd).abs()
    rowmax = Sd.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    print(f"GEMM vs f64 check on {len(rows)} rows: max abs {err.max().item():.3e}, max err/rowmax {(err/rowmax).max().item():.3e}, "
          f"mean signed err/rowmax {((S-Sd)/rowmax).mean().item():.3e}", flush=True)
    print("DONE", flush=True)


if __name__ == "__main__":
    main()

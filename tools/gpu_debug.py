"""Stage-by-stage GPU bring-up check against the CPU oracle (run under gpurun; verbose on purpose)."""
import os
import sys
import time
import ctypes as C
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from mepp_b200 import _lib, synth, layout, motif_layers  # noqa: E402
from mepp_b200.engine import Engine, _ptr  # noqa: E402
from oracle import staging, profile  # noqa: E402


def main():
    print("device", torch.cuda.get_device_name(0), "device_ok", _lib.device_ok(), flush=True)
    cfg = synth.make_config("cfg1", n_seq=int(os.environ.get("DBG_N", 1000)), n_motifs=4,
                            n_perm=int(os.environ.get("DBG_P", 100)))
    ascii_u8, scores, perm = cfg["ascii"], cfg["scores"], cfg["perm_idx"]
    N, L, P = cfg["N"], cfg["L"], cfg["P"]
    margin = cfg["margin"]
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]] + [(cfg["tasks"][0][1], '-')]
    eng = Engine()
    eng.profile_path = 'tensor'        # this tool inspects the tiled operands of the tcgen05 path, stage by stage
    eng.fuse_correlation = False
    eng.stage(ascii_u8, scores, perm, use_gc=True)
    torch.cuda.synchronize()
    codes = staging.ascii_to_codes(ascii_u8)
    # ---- pack
    dec = layout.decode_symbols(eng.sym_T.cpu().numpy(), N, L)
    print("pack codes equal:", np.array_equal(dec, codes), flush=True)
    z_or = staging.gc_ratio_global(codes)
    print("gc equal:", np.array_equal(eng.gc[:N].cpu().numpy(), z_or), flush=True)
    # ---- Y planes
    Y = staging.permuted_scores(scores, perm)
    yh, yl = layout.decode_y_planes(eng.y_planes.cpu().numpy(), eng.C, eng.n_pad, eng.BN, eng.n_tiles_n)
    yq = (yh.astype(np.float64) + yl.astype(np.float64))[:N]
    yc = Y.astype(np.float64) - scores.astype(np.float64).mean()
    sc = (yq[:, 0] / yc[:, 0])[np.abs(yc[:, 0]) > 1e-3]
    print("Y scale (should be one power of two):", sc.min(), sc.max(), flush=True)
    print("Y planes rel err:", np.abs(yq / sc.mean() - yc).max() / np.abs(yc).max(), flush=True)
    # ---- scan
    tables = motif_layers.scan_tables(tasks)
    T = len(tasks)
    batch_dbg = eng.profile(tasks, margin=margin, tables=tables, return_hits=True, return_r=True)
    out = dict(hits=batch_dbg.hits, r=batch_dbg.r, positions=batch_dbg.positions, density=batch_dbg.density)
    torch.cuda.synchronize()
    hits = out['hits']
    m_rows = T * L
    xh, xl = layout.decode_x_planes(eng._bufs['x_planes'].cpu().numpy(), m_rows, eng.n_pad)
    Xg = (xh.astype(np.float64) + xl.astype(np.float64)) / 256.0
    rs = eng._bufs['rowstats'][:m_rows * 3].cpu().numpy().reshape(m_rows, 3)
    zhat = eng.zhat[:N].cpu().numpy().astype(np.float64)
    ors = []
    for t, (pwm, o) in enumerate(tasks):
        orc = profile.profile_task(codes, Y, z_or, pwm, o, margin=margin)
        ors.append(orc)
        X0 = orc['X0']
        h = hits[hits['task'] == t]
        X0g = np.zeros_like(X0)
        X0g[h['seq'], h['pos']] = h['x0']
        nz = int((X0 > 0).sum())
        print(f"task {t} orient {o!r} m={pwm.shape[1]} thr={orc['thr']}: oracle hits {nz}, gpu hits {len(h)}, "
              f"bit-exact: {np.array_equal(X0g.view(np.uint32), X0.view(np.uint32))}", flush=True)
        Xb = profile.blur(X0, margin, dtype=np.float32).astype(np.float64)
        xg = Xg[t * L:(t + 1) * L, :N].T
        print(f"   X planes max abs err {np.abs(xg - Xb).max():.3e} (max {Xb.max():.3f}); pad cols zero: "
              f"{np.all(Xg[t * L:(t + 1) * L, N:] == 0)}", flush=True)
        sx, sxx, sxz = Xb.sum(0), (Xb * Xb).sum(0), (Xb * zhat[:, None]).sum(0)
        print(f"   rowstats err: {np.abs(rs[t*L:(t+1)*L,0]-sx).max():.2e} {np.abs(rs[t*L:(t+1)*L,1]-sxx).max():.2e} "
              f"{np.abs(rs[t*L:(t+1)*L,2]-sxz).max():.2e}", flush=True)
        print(f"   count equal {np.array_equal(out['positions']['count'][t], (X0>0).sum(0))} density equal "
              f"{np.array_equal(out['density'][t,:N], (X0>0).sum(1))}", flush=True)
    # ---- GEMM vs float64 check kernel and vs host (unfused rerun so that S is materialised)
    eng.fuse_correlation = False
    eng.profile(tasks, margin=margin, tables=tables)
    eng.fuse_correlation = True
    S = eng._bufs['S'][:m_rows * eng.ldS].view(m_rows, eng.ldS).cpu().numpy()[:, :eng.C].astype(np.float64)
    Sd = torch.empty((m_rows, eng.C), dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib.mepp_profile_gemm_check(_ptr(eng._bufs['x_planes']), _ptr(eng.y_planes), _ptr(Sd), m_rows,
                                                eng.C, eng.n_pad, None, m_rows, None))
    torch.cuda.synchronize()
    Sc = Sd.cpu().numpy()
    Sh = (Xg * 256.0) @ (yh.astype(np.float64) + yl.astype(np.float64))
    print("check kernel vs host matmul max abs:", np.abs(Sc - Sh).max(), "scale", np.abs(Sh).max(), flush=True)
    err = np.abs(S - Sc)
    print("tcgen05 GEMM vs check: max abs", err.max(), "rel to max", err.max() / max(np.abs(Sc).max(), 1e-30), flush=True)
    if err.max() / max(np.abs(Sc).max(), 1e-30) > 1e-4:
        bad = np.argwhere(err > 1e-4 * np.abs(Sc).max())
        print("  first bad entries (row, col, got, want):", [(int(r), int(c), S[r, c], Sc[r, c]) for r, c in bad[:8]])
        print("  bad rows mod 128:", sorted(set(int(r) % 128 for r, _ in bad))[:40])
        print("  bad cols:", sorted(set(int(c) for _, c in bad))[:40])
    # ---- r and stats
    for t in range(T):
        orc = ors[t]
        rg = out['r'][t]
        d = np.abs(rg.astype(np.float64) - orc['r'].astype(np.float64))
        tol = 1e-5 * np.abs(orc['r']) + 1e-7
        print(f"task {t}: r max abs err {d.max():.3e}, violations of 1e-5 rel+1e-7 abs: {(d > tol).sum()}", flush=True)
    print("DONE", flush=True)


if __name__ == "__main__":
    main()

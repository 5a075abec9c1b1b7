"""Bring-up check of the tensor-core scan stage (run under gpurun): raw int32 accumulators of the first 128 windows
against a numpy emulation of the int8 product, then match lists / r of the tc scan against the CUDA-core lean scan."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine


def quantised(tb):
    """numpy form of tc_weights_kernel for one task table: q[f][j][c] (int), on[f]."""
    m, lut, bias = tb['m'], tb['lut'].astype(np.float64), float(tb['bias'])
    q = np.zeros((2, 32, 4), dtype=np.int64)
    on = [0, 0]
    for f in range(tb['n_filters']):
        w = lut[:m, :, f]
        mx = w.max(axis=1)
        absmax = np.abs(w).max(axis=1).sum()
        delta = 2.0 * (4.0 * m) * absmax * 5.97e-8 + 1e-6
        D = mx.sum() + bias + delta
        if D > 0:
            on[f] = 1
            s = 127.0 / D * (1.0 - 1e-9)
            pen = np.floor(s * (mx[:, None] - w))
            q[f, :m] = -np.minimum(pen, 128).astype(np.int64)
            q[f, 0] += 127
    return q, on


def main():
    case = os.environ.get("TC_CASE", "cfg1")
    n_seq = int(os.environ.get("TC_NSEQ", "256"))
    n_motifs = int(os.environ.get("TC_MOTIFS", "6"))
    cfg = synth.make_config(case, n_seq=n_seq, n_motifs=n_motifs, n_perm=20)
    tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
    L, N = cfg["L"], cfg["N"]
    eng = Engine()
    assert eng.tc_scan
    eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    tables = motif_layers.scan_tables(tasks)
    ld = 2048
    dbg = torch.full((128, ld), -77777, dtype=torch.int32, device='cuda')
    eng._tc_dbg = dbg
    b_tc = eng.profile(tasks, margin=cfg["margin"], return_hits=True, return_r=True, tables=tables)
    torch.cuda.synchronize()
    print("scan used:", eng.last_scan, "paths", eng.last_paths, "fallbacks", eng.fallbacks, flush=True)
    units = eng._last_units_arr
    if units is None:
        units = np.array([(t, -1) for t in range(len(tasks))], dtype=np.int32)
    acc = dbg.cpu().numpy()
    # emulate: flat one-hot stream
    a = cfg["ascii"].reshape(-1)
    oh = np.zeros((a.size + 256, 4), dtype=np.int64)
    for c, ch in enumerate(b"ACGT"):
        oh[:a.size, c] = 4 * (a == ch)
    isn = ~np.isin(a, np.frombuffer(b"ACGT", dtype=np.uint8))
    oh[:a.size][isn] = 1
    flat = oh.reshape(-1)
    # the launch plan of mepp_scan_tc: consecutive units dealt evenly over the ceil(U / 32) N-tiles of one launch
    bad = 0
    col = 0
    U = len(units)
    n_ntiles = (U + 31) // 32
    per = ((U + n_ntiles - 1) // n_ntiles + 3) // 4 * 4
    for u0 in range(0, U, per):
        nu = min(per, U - u0)
        bn = 32 * ((nu + 3) // 4)
        mmax = max(tables[units[u0 + k][0]]['m'] for k in range(nu))
        K = 32 * ((4 * (mmax + 3) + 31) // 32)
        for slot in range(bn // 8):
            if slot < nu:
                tb = tables[units[u0 + slot][0]]
                q, on = quantised(tb)
                m = tb['m']
            for phi in range(4):
                for f in range(2):
                    w = np.zeros(K, dtype=np.int64)
                    if slot < nu and on[f]:
                        for j in range(m):
                            w[4 * (j + phi):4 * (j + phi) + 4] = q[f, j]
                    else:
                        w[0:4] = -1
                    exp = np.array([np.dot(flat[16 * r:16 * r + K], w) for r in range(128)])
                    got = acc[:, col]
                    if not np.array_equal(exp, got):
                        if bad < 6:
                            i = int(np.nonzero(exp != got)[0][0])
                            print(f"ACC MISMATCH col {col} (unit {u0 + slot} phi {phi} f {f}): row {i} expected {exp[i]} got {got[i]}; "
                                  f"{int((exp != got).sum())} rows differ", flush=True)
                        bad += 1
                    col += 1
    print(f"accumulator columns checked: {col}, mismatching columns: {bad}", flush=True)
    # against the CUDA-core lean scan
    eng2 = Engine()
    eng2.tc_scan = False
    eng2.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
    b_ln = eng2.profile(tasks, margin=cfg["margin"], return_hits=True, return_r=True, tables=tables)
    print("reference scan used:", eng2.last_scan, flush=True)
    same_hits = len(b_tc.hits) == len(b_ln.hits) and np.array_equal(b_tc.hits, b_ln.hits)
    print(f"matches tc {len(b_tc.hits)} lean {len(b_ln.hits)} identical: {same_hits}", flush=True)
    if not same_hits:
        A = set(map(tuple, b_tc.hits.tolist())); B = set(map(tuple, b_ln.hits.tolist()))
        print("  missing in tc:", sorted(B - A)[:10]); print("  extra in tc:", sorted(A - B)[:10])
    print("density identical:", np.array_equal(b_tc.density, b_ln.density),
          " r identical:", np.array_equal(b_tc.r, b_ln.r), " max |dr|", float(np.abs(b_tc.r - b_ln.r).max()), flush=True)
    ok = bad == 0 and same_hits
    print("TC CHECK", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

"""Scan kernel alone, repeated, with optional other work in between (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import _lib, synth, motif_layers
from mepp_b200.engine import Engine, _ptr, scan_units

cfg = synth.make_config(os.environ.get("SCALE_CFG", "cfg2"), n_motifs=int(os.environ.get("SCALE_MOTIFS", "24")))
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
eng = Engine()
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
tables = motif_layers.scan_tables(tasks)
T, L, N = len(tables), eng.L, eng.N
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
lut = dev(np.stack([tb['lut'] for tb in tables]).astype(np.float32))
bias = dev(np.array([tb['bias'] for tb in tables], dtype=np.float32))
mlen_h = np.array([tb['m'] for tb in tables], dtype=np.int32)
mlen = dev(mlen_h)
nfilt = dev(np.array([tb['n_filters'] for tb in tables], dtype=np.int32))
qtab = dev(np.stack([tb['qtab'] for tb in tables]).view(np.int32))
qthr = dev(np.stack([tb['qthr'] for tb in tables]).view(np.int32))
units_h = scan_units(tables)
units = dev(units_h)
m_rows = T * L
x = torch.zeros(int(_lib.lib.mepp_x_planes_bytes(m_rows, eng.n_pad)), dtype=torch.uint8, device='cuda')
rowstats = torch.empty((m_rows, 3), dtype=torch.float64, device='cuda')
count = torch.empty((T, L), dtype=torch.int32, device='cuda')
density = torch.empty((T, eng.n_pad), dtype=torch.int32, device='cuda')
hit_count = torch.zeros(1, dtype=torch.int64, device='cuda')
cap = max(1 << 20, int(T * N * L * 2e-3))
hits = torch.empty(cap * 16, dtype=torch.uint8, device='cuda')
junk = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
st = eng._stream()

def scan():
    _lib.check(_lib.lib.mepp_scan(_ptr(eng.sym_T), _ptr(eng.zhat), N, L, T, cfg["margin"], _ptr(lut), _ptr(bias), _ptr(mlen),
                                  _ptr(nfilt), _ptr(qtab), _ptr(qthr), _ptr(units), len(units_h), int(mlen_h.max()),
                                  _ptr(x), _ptr(rowstats), _ptr(count), _ptr(density), _ptr(hits), cap, _ptr(hit_count), None, 0, None, None, None, st), "scan")

def reset():
    _lib.check(_lib.lib.mepp_x_planes_reset(_ptr(x), _ptr(hits), cap, _ptr(hit_count), N, L, T, cfg["margin"], st), "reset")

def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)

mode = os.environ.get("SCANB_MODE", "plain")
small = tasks[:8]
for i in range(8):
    t = timed(scan)
    extra = ""
    if mode == "gemm":
        eng.profile(small, margin=cfg["margin"]); torch.cuda.synchronize()
    if mode in ("reset", "all"):
        extra += f" reset {timed(reset):.3f}"
    if mode in ("junk", "all"):
        extra += f" junk {timed(lambda: junk.fill_(i)):.3f}"
    print(f"{mode} scan {i}: {t:.3f} ms  matches {int(hit_count.item())}{extra}", flush=True)

if os.environ.get("SCANB_DIFF"):
    HIT = np.dtype([('task', '<i4'), ('seq', '<i4'), ('pos', '<i4'), ('x0', '<f4')])
    def grab():
        n = int(hit_count.item())
        h = hits[:n * 16].cpu().numpy().view(HIT).copy()
        return h[np.lexsort((h['pos'], h['seq'], h['task']))]
    scan(); torch.cuda.synchronize(); ref = grab()
    for i in range(6):
        reset(); scan(); torch.cuda.synchronize()
        h = grab()
        if len(h) != len(ref) or not np.array_equal(h, ref):
            a = set(map(tuple, ref.tolist())); b = set(map(tuple, h.tolist()))
            miss = sorted(a - b); extra = sorted(b - a)
            print(f"run {i}: {len(h)} vs {len(ref)}; missing {len(miss)} extra {len(extra)}")
            for rec in miss[:12]:
                print("   missing", rec, "kb", rec[1] // 32, "lane", rec[1] % 32, "chunk", rec[2] // 32, "m", int(mlen_h[rec[0]]))
            for rec in extra[:6]:
                print("   extra  ", rec)
        else:
            print(f"run {i}: identical")

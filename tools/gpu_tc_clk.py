"""Cycle stamps of CTA 0's MMA / epilogue pipeline in the tensor-core scan (run under gpurun; tuning aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mepp_b200 import synth, motif_layers
from mepp_b200.engine import Engine

cfg = synth.make_config("cfg2", n_motifs=41)
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]]
eng = Engine()
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"], use_gc=True)
tables = motif_layers.scan_tables(tasks)
clk = torch.zeros((10, 256), dtype=torch.int64, device='cuda')


class Dbg:      # quacks like the accumulator dump: shape[1] < 0 selects the stamp mode of mepp_scan_tc
    shape = (5, -1)
    def data_ptr(self): return clk.data_ptr()


eng.profile(tasks, margin=2, tables=tables)          # warm
eng._tc_dbg = Dbg()
eng.profile(tasks, margin=2, tables=tables)
torch.cuda.synchronize()
c = clk.cpu().numpy()
t0 = c[0][0]
names = ["mma: tempty ok", "mma: committed", "epi: tfull ok", "epi: released", "epi: tested", "mma: tile there", "mma: 1st issued",
         "mma: all issued", "-", "mma: loop top"]
n = int(os.environ.get("CLK_N", "40"))
print("n-tile " + " ".join(f"{names[k]:>16s}" for k in (9, 5, 0, 6, 7, 1, 2, 3, 4)))
for i in range(n):
    print(f"{i:6d} " + " ".join(f"{int(c[k][i] - t0) if c[k][i] else 0:16d}" for k in (9, 5, 0, 6, 7, 1, 2, 3, 4)))
d = np.diff(c[1][8:200])
print("steady-state cycles per N-tile (mma commit to commit): median", np.median(d), "mean", d.mean())
print("commit -> epilogue sees tfull: median", np.median(c[2][8:200] - c[1][8:200]))
print("epilogue release -> mma sees tempty (same stage, +4): median", np.median(c[0][12:204] - c[3][8:200]))
print("epilogue: tfull -> released", np.median(c[3][8:200] - c[2][8:200]), " released -> tested", np.median(c[4][8:200] - c[3][8:200]))

"""torchrun check: staging by NCCL broadcast gives the same results as staging from each rank's own host copy."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from mepp_b200 import synth, parallel
from mepp_b200.engine import Engine

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
cfg = synth.make_config("cfg1", n_seq=600, n_perm=50)
tasks = [(pwm, o) for _, pwm, o in cfg["tasks"]][:8]
eng = Engine(device=local)
eng.stage(cfg["ascii"], cfg["scores"], cfg["perm_idx"])
a = eng.profile(tasks, margin=2, return_r=True, return_hits=True)
args = (cfg["ascii"], cfg["scores"], cfg["perm_idx"]) if rank == 0 else (None, None, None)
eng.stage_broadcast(*args, src=0)
b = eng.profile(tasks, margin=2, return_r=True, return_hits=True)
ok = np.array_equal(a.hits, b.hits) and np.allclose(a.r, b.r, rtol=1e-6, atol=1e-9) and np.array_equal(a.density, b.density)
print(f"rank {rank}: broadcast staging equal: {ok}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

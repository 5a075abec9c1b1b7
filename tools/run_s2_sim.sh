#!/bin/bash
# per-rank load of an N-GPU strong-scaling run on one GPU, with the host phases of a profile() call
for SW in "$@"; do
timeout 600 python bench.py --steps 20 --warmup 3 --simulate-world $SW --no-cpu-baseline --no-accuracy --no-clocks --no-tensor-leg > gpurun_out/s2_sim$SW.json 2>gpurun_out/s2_sim$SW.err
python - gpurun_out/s2_sim$SW.json $SW <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("sim", sys.argv[2], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()}, d["details"]["host_ms_per_step"], "launches", d["gpu_launches"])
PY
done

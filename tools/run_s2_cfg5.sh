#!/bin/bash
# GPU tests, then cfg5 at full size (1e6 x 200 bp, P = 1e4: the permutation index matrix is staged in chunks)
SECONDS=0
timeout 1500 python bench.py --workload cfg5 --steps ${1:-3} --warmup 3 --no-cpu-baseline --no-fullrun --no-clocks --no-tensor-leg > gpurun_out/s2_cfg5.json 2>gpurun_out/s2_cfg5.err; echo rc=$?
grep -v "^\s" gpurun_out/s2_cfg5.err | tail -5
python - gpurun_out/s2_cfg5.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2),
      {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
print("accuracy", d.get("accuracy"))
PY
nvidia-smi --query-gpu=memory.used --format=csv
echo elapsed $SECONDS s

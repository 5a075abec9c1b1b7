#!/bin/bash
# session baseline: GPU tests, cfg2 and cfg3 stage times
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2_gputest.log; cat gpurun_out/s2_gputest.log
for W in cfg2 cfg3; do
timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-fullrun --no-tensor-leg > gpurun_out/s2_base_$W.json 2>gpurun_out/s2_base_$W.err
python - gpurun_out/s2_base_$W.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
PY
done

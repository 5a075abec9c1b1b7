#!/bin/bash
W=${1:-cfg3}
for M in 0 4 6 12 14 15; do
  MEPP_TC_MODE=$M timeout 200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/mode$M.json 2>gpurun_out/mode$M.err
  python - gpurun_out/mode$M.json $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("mode", sys.argv[2], "scan ms", round(d["kernels"]["scan"]["ms_per_step"], 3))
PY
done

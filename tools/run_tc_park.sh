#!/bin/bash
for P in 0 200 1000 4000; do
  MEPP_TC_PARK_NS=$P timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/park$P.json 2>gpurun_out/park$P.err
  python - gpurun_out/park$P.json $P <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("park", sys.argv[2], "scan ms", round(d["kernels"]["scan"]["ms_per_step"], 3))
PY
done
MEPP_TC_PARK_NS=0 CLK_N=8 timeout 120 python tools/gpu_tc_clk.py 2>&1 | tail -13

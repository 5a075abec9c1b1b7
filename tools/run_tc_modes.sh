#!/bin/bash
# timing experiments of the tensor-core scan (MEPP_TC_MODE bits: 1 no TMEM reads / epilogue work, 2 no MMAs,
# 4 TMEM reads but no sign tests, 8 no stream copies): tools/run_tc_modes.sh [workload] [modes...]
W=${1:-cfg3}; shift
for M in ${@:-0 1 2 4 5}; do
  MEPP_TC_MODE=$M timeout 200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-clocks --no-accuracy --no-fullrun --no-tensor-leg > gpurun_out/mode$M.json 2>gpurun_out/mode$M.err
  python - gpurun_out/mode$M.json $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("mode", sys.argv[2], "scan ms", round(d["kernels"]["scan"]["ms_per_step"], 3))
PY
done

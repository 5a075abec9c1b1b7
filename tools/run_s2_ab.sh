#!/bin/bash
# A/B of an environment switch on the cfg3 / cfg2 step: tools/run_s2_ab.sh VAR valueA valueB [workload]
VAR=$1; A=$2; B=$3; W=${4:-cfg3}
for V in $A $B $A $B; do
env $VAR=$V timeout 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline --no-accuracy --no-clocks --no-fullrun --no-tensor-leg > gpurun_out/s2_ab_$V.json 2>gpurun_out/s2_ab_$V.err
python - gpurun_out/s2_ab_$V.json $VAR=$V <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
PY
done

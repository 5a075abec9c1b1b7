#!/bin/bash
# bring-up checks of the tensor-core scan + pipeline cycle stamps + cfg2 / cfg3 stage times (run under gpurun)
timeout 120 python tools/gpu_tc_check.py > gpurun_out/tc_check.log 2>&1; echo check rc=$?
tail -3 gpurun_out/tc_check.log
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/tc_check2.log 2>&1; rc=$?; echo check2 rc=$rc
tail -5 gpurun_out/tc_check2.log
if [ $rc -ne 0 ]; then exit 1; fi
CLK_N=10 timeout 120 python tools/gpu_tc_clk.py > gpurun_out/tc_clk.log 2>&1; tail -15 gpurun_out/tc_clk.log
for W in cfg2 cfg3; do
timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/tc_ab_$W.json 2>gpurun_out/tc_ab_$W.err
python - gpurun_out/tc_ab_$W.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
PY
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

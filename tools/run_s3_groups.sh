#!/bin/bash
# parity suite with the default kernels, then the cfg3 step for 4 / 3 / 2 / 1 task groups (MEPP_GROUPS)
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/s3_check.log 2>&1; echo check rc=$?
tail -3 gpurun_out/s3_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; rc=$?; echo pytest rc=$rc
tail -3 gpurun_out/s3_pytest.log
for G in ${@:-4 3 2 1}; do
MEPP_GROUPS=$G timeout 600 python bench.py --workload cfg3 --steps 30 --warmup 3 --no-cpu-baseline --no-accuracy --no-clocks --no-fullrun --no-tensor-leg > gpurun_out/s3_g$G.json 2>gpurun_out/s3_g$G.err
python - gpurun_out/s3_g$G.json groups=$G <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
PY
done

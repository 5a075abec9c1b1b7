#!/bin/bash
# stage times of a cfg2 step + ncu --set full of the tensor-core scan kernel (run under gpurun)
TAG=${1:-r02_tc}
python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_err.log
python - gpurun_out/${TAG}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
PY
python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_scan_kernel" -s 4 -c 2 -o gpurun_out/${TAG}_scan \
  python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c 1-200

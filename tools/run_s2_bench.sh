#!/bin/bash
# one bench line per call: tools/run_s2_bench.sh <tag> <bench args...>
TAG=$1; shift
timeout 900 python bench.py "$@" > gpurun_out/${TAG}.json 2>gpurun_out/${TAG}.err; echo rc=$?
tail -3 gpurun_out/${TAG}.err
python - gpurun_out/${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2),
      {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
print("accuracy", d.get("accuracy"))
print("cpu", d.get("cpu_baseline"))
PY

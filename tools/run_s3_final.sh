#!/bin/bash
# round-2 closing run: GPU test suite, the default bench line, the ncu launch list of the same command (2 steps) and one
# ncu --set full capture of tc_scan4i_kernel
T=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err; echo bench rc=$?
python - gpurun_out/${T}_bench_cfg3.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2),
      {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
print("roofline", d.get("roofline"))
print("accuracy", d.get("accuracy"))
print("cpu", d.get("cpu_baseline"))
print("fullrun", d.get("fullrun"))
print("clocks", d.get("clocks"))
PY
ARGS="--workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-clocks --no-accuracy --no-fullrun --no-tensor-leg"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches_cfg3.csv \
  python bench.py $ARGS > gpurun_out/${T}_ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:tc_scan4i -s 5 -c 1 -f -o gpurun_out/${T}_tc_scan4i \
  python bench.py $ARGS > gpurun_out/${T}_ncu_tc.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/${T}_*

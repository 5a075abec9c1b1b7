#!/bin/bash
# strong-scaling bench line at N GPUs: tools/run_s2_multi.sh <N> <tag> <bench args...>
N=$1; TAG=$2; shift; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N "$@" > gpurun_out/${TAG}.json 2>gpurun_out/${TAG}.err; echo rc=$?
grep -v "^\[W\|^W\|^\*\*\*\|^$" gpurun_out/${TAG}.err | tail -5
python - gpurun_out/${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], d["n_gpus"], d["scaling"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2),
      {k: round(v.get("ms_per_step") or v.get("ms") or 0, 3) for k, v in d["kernels"].items()})
print("accuracy", {k: d["accuracy"][k] for k in ("match_sets_equal", "r_within_1e5_gate", "r_max_rel_err", "pvalue_flips")} if d.get("accuracy") else None)
PY

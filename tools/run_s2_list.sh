#!/bin/bash
# per-kernel device time of a step (ncu launch list): tools/run_s2_list.sh <tag> [workload]
TAG=${1:-r02l}; W=${2:-cfg3}
ARGS="--workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-clocks --no-accuracy --no-fullrun --no-tensor-leg"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py $ARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
python - gpurun_out/${TAG}_launches.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[k].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[v].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{t/1e3:10.1f} us {c:5d} launches {t/1e3/c:9.1f} us each {100*t/tot:5.1f}%  {n}")
PY

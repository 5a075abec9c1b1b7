#!/bin/bash
# launch list (per-kernel device time) of one bench step: tools/run_launch_list.sh <workload> <tag>
W=${1:-cfg2}; TAG=${2:-r02}
python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_ncu.log 2>&1
python - gpurun_out/${TAG}_launches.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[k].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[v].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t/1e3:10.1f} us {c:5d} launches {100*t/tot:5.1f}%  {n}")
PY

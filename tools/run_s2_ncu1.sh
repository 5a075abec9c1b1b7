#!/bin/bash
# ncu --set full of ONE kernel of a cfg3 step: tools/run_s2_ncu1.sh <tag> <kernel regex> [skip] [workload]
TAG=$1; K=$2; SKIP=${3:-5}; W=${4:-cfg3}
ARGS="--workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-clocks --no-accuracy --no-fullrun --no-tensor-leg"
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu.log 2>&1
echo rc=$?; ls -la gpurun_out/${TAG}.ncu-rep

#!/bin/bash
# GPU tests + full-run wall clock (run_mepp: files in, directory out) on cfg2 and cfg3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for W in "$@"; do
timeout 900 python bench.py --mode fullrun --workload $W > gpurun_out/s2_fullrun_$W.json 2>gpurun_out/s2_fullrun_$W.err; echo rc=$?
tail -3 gpurun_out/s2_fullrun_$W.err
cat gpurun_out/s2_fullrun_$W.json | cut -c 1-1500
done

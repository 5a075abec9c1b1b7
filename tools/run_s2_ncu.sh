#!/bin/bash
# launch list + ncu --set full of the main kernels of one step (run under gpurun): tools/run_s2_ncu.sh <tag> <workload> [kernels...]
TAG=${1:-r02b}; W=${2:-cfg3}; shift; shift
ARGS="--workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-clocks --no-accuracy --no-fullrun --no-tensor-leg"
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py $ARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
for K in ${@:-tc_scan_kernel match_gather_kernel match_blur_kernel row_select_kernel pooled_pass_kernel bucket_entries_kernel tc_exact_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 5 -c 1 -f -o gpurun_out/${TAG}_$K \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_$K.log 2>&1
  echo $K rc=$?
done
ls -la gpurun_out/${TAG}_*.ncu-rep

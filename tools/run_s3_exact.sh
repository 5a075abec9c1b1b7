#!/bin/bash
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/s3_check.log 2>&1; echo check rc=$?
tail -3 gpurun_out/s3_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_exact_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/s3_exact_pytest.log
tools/run_s2_bench.sh s3_exact --workload cfg3 --steps 30 --warmup 3 --no-cpu-baseline --no-clocks --no-fullrun --no-tensor-leg | head -4 | cut -c1-900

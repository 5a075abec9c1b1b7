#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "perm_stats or golden or finalisation or config1" > gpurun_out/s3_k3_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/s3_k3_pytest.log
tools/run_s2_bench.sh s3_k3 --workload cfg3 --steps 30 --warmup 3 --no-cpu-baseline --no-accuracy --no-clocks --no-fullrun --no-tensor-leg | head -3

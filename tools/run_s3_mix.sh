#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "match_driven or lean_scan or config1 or full_size" > gpurun_out/s3_mix_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/s3_mix_pytest.log
MEPP_GATHER_MIX=1 timeout 600 python -m pytest tests -m gpu -x -q -k "match_driven or config1 or full_size" > gpurun_out/s3_mix_pytest2.log 2>&1; echo pytest-mix rc=$?; tail -2 gpurun_out/s3_mix_pytest2.log
tools/run_s2_ab.sh MEPP_GATHER_MIX 0 1 cfg3

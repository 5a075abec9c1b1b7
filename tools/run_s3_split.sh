#!/bin/bash
# half-stage form of tc_scan4i_kernel (MEPP_TC_SPLIT=2): parity suite, A/B against whole stages, timing modes
export MEPP_TC_SPLIT=2
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/s3_split_check.log 2>&1; echo check rc=$?
tail -4 gpurun_out/s3_split_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_split_pytest.log 2>&1; rc=$?; echo pytest rc=$rc
tail -4 gpurun_out/s3_split_pytest.log
unset MEPP_TC_SPLIT
tools/run_s2_ab.sh MEPP_TC_SPLIT 1 2 ${1:-cfg3}
MEPP_TC_SPLIT=2 tools/run_tc_modes.sh cfg3 4 1 3

#!/bin/bash
# one issuer per accumulator stage (tc_scan4i_kernel, MEPP_TC_ISSUERS=4): parity suite, then A/B against two issuers
export MEPP_TC_ISSUERS=4
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/s3_i4_check.log 2>&1; echo check rc=$?
tail -4 gpurun_out/s3_i4_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_i4_pytest.log 2>&1; rc=$?; echo pytest rc=$rc
tail -4 gpurun_out/s3_i4_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
unset MEPP_TC_ISSUERS
tools/run_s2_ab.sh MEPP_TC_ISSUERS 2 4 ${1:-cfg3}

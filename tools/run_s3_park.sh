#!/bin/bash
# tc_scan4i_kernel: parity check, then the scan time with suspend-time hints on the epilogue's barrier waits
TC_CASE=cfg2 TC_NSEQ=3000 TC_MOTIFS=41 timeout 200 python tools/gpu_tc_check.py > gpurun_out/s3_check.log 2>&1; echo check rc=$?
tail -3 gpurun_out/s3_check.log
for P in 0 200 600 2000; do echo park $P; MEPP_TC_PARK_NS=$P tools/run_tc_modes.sh cfg3 0; done

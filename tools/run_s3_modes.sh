#!/bin/bash
# tc_scan4i_kernel: timing modes (1 no epilogue reads, 2 no products, 4 reads without tests, 8 half of the reads), park hints
export MEPP_TC_ISSUERS=4
tools/run_tc_modes.sh cfg3 ${@:-0 4 8 12 1}
for P in 200 1000; do echo park $P; MEPP_TC_PARK_NS=$P tools/run_tc_modes.sh cfg3 0; done

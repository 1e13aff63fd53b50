#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/trace_*.txt
C5_TRACE_FILE=gpurun_out/trace_band_l1.txt timeout 600 python scripts/exp_lanes.py C3 --rows "430,555" --lanes 1 --views 6 > /dev/null 2>&1
python scripts/trace_blocks.py gpurun_out/trace_band_l1.txt | head -12
C5_TRACE_FILE=gpurun_out/trace_band_l2.txt timeout 600 python scripts/exp_lanes.py C3 --rows "430,555" --lanes 2 --views 8 > /dev/null 2>&1
python scripts/trace_blocks.py gpurun_out/trace_band_l2.txt | head -40
gzip -f gpurun_out/trace_band_l1.txt gpurun_out/trace_band_l2.txt
exit 0

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== normal"
timeout 600 python scripts/exp_lanes.py C3 --rows "430,555;800,925" --lanes 1,2 2>&1 | tee gpurun_out/exp_lanes_prep.jsonl | cut -c1-300
echo "== per-view preparation skipped (same view repeated)"
C5_SKIP_PREP=1 timeout 600 python scripts/exp_lanes.py C3 --rows "430,555;800,925;0,0" --lanes 1,2 2>&1 | tee -a gpurun_out/exp_lanes_prep.jsonl | cut -c1-300
echo "== grazing kernel serial after the pixel kernel"
C5_GRAZE_SERIAL=1 timeout 600 python scripts/exp_lanes.py C3 --rows "430,555;0,0" --lanes 1,2 2>&1 | tee -a gpurun_out/exp_lanes_prep.jsonl | cut -c1-300
exit 0

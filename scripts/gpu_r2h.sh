#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;430,555;800,925;430,680" --lanes 1,2 2>&1 | tee gpurun_out/exp_lanes.jsonl | cut -c1-300
echo "== ncu: solid_mask (whole image)"
CMD="python scripts/exp_configs.py C3 --top 0 --reps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solid_mask -s 4 -c 2 -f -o gpurun_out/mask_full $CMD > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
exit 0

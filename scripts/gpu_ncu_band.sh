#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python scripts/exp_configs.py C3 --top 0 --reps 1 --no-solids --rows 395,405"
timeout 600 $CMD > gpurun_out/plain_band.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:tet_walk_fp64 -s 2 -c 1 -f -o gpurun_out/walk_band $CMD > gpurun_out/ncu_band.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_band.log | cut -c1-300
exit 0

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;430,555;800,925" --lanes 1,2,3,4 2>&1 | tee gpurun_out/exp_lanes.jsonl | cut -c1-300
timeout 600 python scripts/exp_lanes.py C3 --view 0.4,0.3 --rows "0,0;800,925" --lanes 1,2,3 2>&1 | tee -a gpurun_out/exp_lanes.jsonl | cut -c1-300
exit 0

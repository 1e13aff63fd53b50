#!/bin/bash
# Round 2, first GPU call (ONE GPU): parity suite, step records ("rec") against cells, views in flight
# on a one-eighth band with the grazing-ray kernel after the pixel kernel (what every rank of an
# 8-GPU run does).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["variant"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"], "graze", d["grazing_rays"])
    else: print(l.rstrip())
'
echo "== step records (rec) against cells (default)"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --variants default,rec --rows "0,1800;430,555" 2>&1 | tee gpurun_out/exp_rec.jsonl | python -c "$show"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --variants default,rec --view 0.4,0.3 2>&1 | tee -a gpurun_out/exp_rec.jsonl | python -c "$show"
echo "== views in flight, grazing-ray kernel after the pixel kernel"
C5_GRAZE_SERIAL=1 timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;430,555;800,925" --lanes 1,2,3,4 --views 24 2>&1 | tee gpurun_out/exp_lanes_serial.jsonl | cut -c1-250
echo "== views in flight, grazing-ray kernel beside"
timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;430,555" --lanes 1,2,3 --views 24 2>&1 | tee gpurun_out/exp_lanes_beside.jsonl | cut -c1-250
timeout 900 python scripts/exp_configs.py C5 --top 0 --reps 3 --variants default,rec --res 2400,1800 2>&1 | tee -a gpurun_out/exp_rec.jsonl | python -c "$show"
exit 0

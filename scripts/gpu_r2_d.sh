#!/bin/bash
# ONE GPU: the tile-skipping solid mask (parity first), then views in flight / prologue priority on
# the bands an 8-GPU run cuts.
set -u
mkdir -p gpurun_out
echo "== pytest (mask, goldens, reference configs)" && timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mask or golden or reference or c1 or pinned or submit" > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -4 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], "walk", d["ms_walk"], "graze", d["ms_graze"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"])
    else: print(l.rstrip())
'
timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911;702,828;534,702" 2>&1 | tee gpurun_out/exp_mask_tiles.jsonl | python -c "$show"
timeout 600 python scripts/exp_configs.py C3 --reps 5 --view 0.4,0.3 2>&1 | tee -a gpurun_out/exp_mask_tiles.jsonl | python -c "$show"
for dbg in "prep_priority=0" "prep_priority=1"; do
  echo "== lanes, $dbg"
  timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;534,702;828,911" --lanes 4,6,8 --views 32 --debug $dbg 2>&1 | tee -a gpurun_out/exp_lanes_prep.jsonl | cut -c1-260
done
exit 0

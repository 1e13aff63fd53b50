#!/bin/bash
# Quick single-GPU check after a kernel change: parity tests, row-band timings, whole-view timings.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --variants default --rows "0,1800;0,400;430,555;800,925;395,405" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['variant'], d['rows'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'steps M', round(d['tet_steps']/1e6,1), 'G/s', d['walk_Gsteps_per_s'])
"
timeout 900 python scripts/exp_configs.py C3 --variants default --top 0 --precision 64,32 2>&1 | tee gpurun_out/exp_c3.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['view'], d['precision'], d['variant'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'G/s', d['walk_Gsteps_per_s'])
"
timeout 900 python scripts/exp_configs.py C3 --view 0.4,0.3 --variants default --top 0 --precision 64,32 2>&1 | tee -a gpurun_out/exp_c3.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['view'], d['precision'], d['variant'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'G/s', d['walk_Gsteps_per_s'])
"
timeout 1200 python scripts/exp_configs.py C1 C2 C5 --variants default --top 0 2>&1 | tee gpurun_out/exp_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['variant'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'total', d['ms_total'], 'G/s', d['walk_Gsteps_per_s'])
"
exit 0

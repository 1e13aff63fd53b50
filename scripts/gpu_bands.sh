#!/bin/bash
# Where does a row band's time go? Same view, different bands of the C3 image on one GPU.
set -u
mkdir -p gpurun_out
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --rows "0,1800;0,400;400,430;430,555;800,925;1275,1400;1370,1400;1400,1800;395,405" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['rows'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'steps M', round(d['tet_steps']/1e6,1), 'G/s', d['walk_Gsteps_per_s'])
"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --view 0.4,0.3 --rows "0,1800;800,925;800,1025" 2>&1 | tee -a gpurun_out/exp_bands.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('oblique', d['rows'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'steps M', round(d['tet_steps']/1e6,1), 'G/s', d['walk_Gsteps_per_s'])
"
echo "== compute-sanitizer memcheck on the smoke test"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
exit 0

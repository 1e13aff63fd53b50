#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest parity" && timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
for B in 1000000 128 64 32; do
  echo "== query budget $B"
  C5_QUERY_BUDGET=$B timeout 600 python scripts/exp_lanes.py C3 --rows "430,555;0,0" --lanes 1,2 --views 12 2>&1 | cut -c1-300
done
C5_QUERY_BUDGET=64 timeout 600 python scripts/exp_lanes.py C3 --view 0.4,0.3 --rows "800,925;0,0" --lanes 1,2 --views 12 2>&1 | cut -c1-300
C5_QUERY_BUDGET=1000000 timeout 600 python scripts/exp_lanes.py C3 --view 0.4,0.3 --rows "0,0" --lanes 2 --views 12 2>&1 | cut -c1-300
rm -f gpurun_out/trace_*.txt
C5_TRACE_FILE=gpurun_out/trace_band_b64.txt timeout 600 python scripts/exp_lanes.py C3 --rows "430,555" --lanes 1 --views 4 > /dev/null 2>&1
python scripts/trace_blocks.py gpurun_out/trace_band_b64.txt 2>/dev/null | head -4 | cut -c1-400
timeout 600 python scripts/exp_configs.py C3 --top 0 --reps 3 --rows "0,1800;430,555" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['rows'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'total', d['ms_total'], 'graze', d['grazing_rays'])
"
exit 0

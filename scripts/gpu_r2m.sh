#!/bin/bash
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["view"], d["variant"], d["precision"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "graze", d["grazing_rays"])
'
echo "== pytest parity" && timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
echo "== previous build"; C5GPU_LIBRARY=$PWD/build/libc5gpu_prev.so timeout 600 python scripts/exp_configs.py C3 --top 0 --reps 3 --rows "0,1800;430,555" 2>&1 | python -c "$show"
for B in 100000000 128 64; do
echo "== this build, budget $B"; C5_QUERY_BUDGET=$B timeout 600 python scripts/exp_configs.py C3 --top 0 --reps 3 --variants default,r80 --rows "0,1800;430,555" 2>&1 | python -c "$show"
done
echo "== oblique, budget 64 / off; fp32"
C5_QUERY_BUDGET=64 timeout 600 python scripts/exp_configs.py C3 --top 0 --reps 3 --view 0.4,0.3 --precision 64,32 2>&1 | python -c "$show"
C5_QUERY_BUDGET=100000000 timeout 600 python scripts/exp_configs.py C3 --top 0 --reps 3 --view 0.4,0.3 2>&1 | python -c "$show"
rm -f gpurun_out/trace_*.txt
C5_TRACE_FILE=gpurun_out/trace_band_b64.txt timeout 600 python scripts/exp_lanes.py C3 --rows "430,555" --lanes 1,2 --views 4 2>&1 | cut -c1-200
python scripts/trace_blocks.py gpurun_out/trace_band_b64.txt 2>/dev/null | head -3 | cut -c1-400
exit 0

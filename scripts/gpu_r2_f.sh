#!/bin/bash
# ONE GPU: full parity suite (with the C4 sweep test), mask timings, ncu --set full of the mask kernels, bench.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -12 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; }
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], "walk", d["ms_walk"], "graze", d["ms_graze"], "mask", d["ms_mask"], "total", d["ms_total"])
    else: print(l.rstrip())
'
timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911;702,828" 2>&1 | tee gpurun_out/exp_mask_b.jsonl | python -c "$show"
echo "== ncu full: mask kernels, whole view"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solid_mask" -s 10 -c 4 -f -o gpurun_out/mask_r02 python scripts/exp_configs.py C3 --reps 1 > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
ncu -i gpurun_out/mask_r02.ncu-rep --page raw --csv > gpurun_out/mask_r02_raw.csv 2>/dev/null
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?
cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
exit 0

#!/bin/bash
# ONE GPU: warp-cooperative solid-mask passes + cached static-solid footprint: parity, timings against the
# per-face kernels (experiments build), ncu of the mask kernels, bench with the clock poller primed.
set -u
mkdir -p gpurun_out
echo "== pytest (mask-related first)" && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mask or golden or c1 or c2 or static or row_bands or sibling" > gpurun_out/pytest_gpu_mask.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu_mask.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu_mask.log | cut -c1-300; exit $rc; }
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], d.get("debug"), "mask", d["ms_mask"], "walk", d["ms_walk"], "total", d["ms_total"])
    else: print(l.rstrip())
'
ROWS="0,1800;828,911;702,828;534,702"
rm -f gpurun_out/exp_mask_coop.jsonl
echo "== product library: cooperative passes, static footprint cached"
timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "$ROWS" 2>&1 | tee -a gpurun_out/exp_mask_coop.jsonl | python -c "$show"
echo "== product library: cooperative passes, static footprint NOT cached"
timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "$ROWS" --debug no_static_mask=1 2>&1 | tee -a gpurun_out/exp_mask_coop.jsonl | python -c "$show"
echo "== experiments build: per-face passes, not cached (= the previous kernels)"
C5GPU_LIBRARY=build/exp/libc5gpu_exp.so timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "$ROWS" --debug no_static_mask=1,mask_per_face=1 2>&1 | tee -a gpurun_out/exp_mask_coop.jsonl | python -c "$show"
echo "== oblique view, product"
timeout 600 python scripts/exp_configs.py C3 --reps 5 --view 0.4,0.3 2>&1 | tee -a gpurun_out/exp_mask_coop.jsonl | python -c "$show"
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 --timeline gpurun_out/timeline_n1.json > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?
cut -c1-700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu full: mask kernels, whole view"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solid_mask|mask_tile" -s 12 -c 3 -f -o gpurun_out/mask_r02_coop python scripts/exp_configs.py C3 --reps 1 > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
ncu -i gpurun_out/mask_r02_coop.ncu-rep --page raw --csv > gpurun_out/mask_r02_coop_raw.csv 2>/dev/null
echo "== rest of the parity suite"
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -12 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; }
exit 0

#!/bin/bash
# The first GPU call of the next round (ONE GPU, ~3 GPU-minutes): everything that was written after the
# GPU budget of round 1 ran out. Stops at the first failing stage.
#   1. parity, including the two tests that never ran on a GPU (search budget, step-record variant)
#   2. step records against cells: whole C3 view, a one-wave band, the oblique view, C5
#   3. bench.py as the driver runs it
# Then, separately and in this order (N-GPU box time is charged N times; each bench has its own watchdog):
#   gpurun --gpus 2 -- bash scripts/gpu_multi.sh     # two-process image test, sendrecv / p2p at N = 2
#   gpurun --gpus 8 -- bash scripts/gpu_scale.sh     # N = 1, 2, 4, 8 the way the driver launches them (~1.5 min of box)
set -u
mkdir -p gpurun_out
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
timeout 900 python scripts/exp_configs.py C5 --top 0 --reps 3 --variants default,rec 2>&1 | tee -a gpurun_out/exp_rec.jsonl | python -c "$show"
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?
cut -c1-3000 gpurun_out/bench.json; tail -4 gpurun_out/bench.err
exit $rc

#!/bin/bash
# L2 slab prefetch: parity, then lookahead 0/1/2 on the whole C3 image, on row bands, oblique, C5.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["precision"], d["variant"], "pf", d["prefetch"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"], "graze", d["grazing_rays"])
    else: print(l.rstrip())
'
echo "== C3 README view: whole image and bands, prefetch 0/1/2"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --prefetch 0,1,2 --rows "0,1800;430,555;800,925;0,400;675,1125" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "$show"
echo "== oblique, fp32"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --prefetch 0,1,2 --view 0.4,0.3 --precision 64,32 2>&1 | tee gpurun_out/exp_c3.jsonl | python -c "$show"
echo "== other configs"
timeout 1500 python scripts/exp_configs.py C1 C2 C5 --top 0 --prefetch 0,1 2>&1 | tee gpurun_out/exp_configs.jsonl | python -c "$show"
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-1800 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
exit 0
